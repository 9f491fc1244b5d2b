#!/bin/bash
# round 2, trip E (1 GPU): tests, then forward tail split A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -8 | tee gpurun_out/r2e_pytest.log
echo "== bench A/B (forward tail split on/off, interleaved)"
for v in 1 0 1 0; do
  VV_FWD_TAIL=$v timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r2e_bench_$v.json 2> gpurun_out/r2e_bench_$v.err
  tail -1 gpurun_out/r2e_bench_$v.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_bench_$v.json").read().strip().splitlines()[-1])
    print("fwd_tail=$v:", round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"],
          {k:(round(v["ms"],4), round(v["frac"],3) if v.get("frac") else None) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"], "loss", d["loss"])
except Exception as e:
    print("no result", e)
PY
done
VV_FWD_TAIL=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 tail=1', round(d['value']), d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()})"
VV_FWD_TAIL=0 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 tail=0', round(d['value']), d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()})"
