#!/bin/bash
# round 2, trip D (1 GPU): tests, then fused-finish (drain) A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -12 | tee gpurun_out/r2d_pytest.log
echo "== bench A/B (fused update on/off, twice each, interleaved)"
for v in 1 0 1 0; do
  VV_FUSED_UPDATE=$v timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r2d_bench_$v.json 2> gpurun_out/r2d_bench_$v.err
  tail -1 gpurun_out/r2d_bench_$v.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_bench_$v.json").read().strip().splitlines()[-1])
    print("fused_update=$v:", round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"],
          {k:(round(v["ms"],4), round(v["frac"],3) if v.get("frac") else None) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"], "loss", d["loss"])
except Exception as e:
    print("no result", e)
PY
done
