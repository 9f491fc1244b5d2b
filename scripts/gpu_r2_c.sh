#!/bin/bash
# round 2, trip C (1 GPU): rank kernel v2.1 A/B + tests + profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== rank kernel alone (public API: atomics path)"
VV_RANK_V2=0 timeout 120 python scripts/bench_rank.py 0 1 2>&1 | tail -1
VV_RANK_V2=1 timeout 120 python scripts/bench_rank.py 0 1 2>&1 | tail -1
echo "== tests"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -12 | tee gpurun_out/r2c_pytest.log
echo "== bench A/B"
for v in "1 0" "0 0"; do
  set -- $v
  VV_RANK_V2=$1 VV_FUSED_UPDATE=$2 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r2c_bench_$1$2.json 2> gpurun_out/r2c_bench_$1$2.err
  tail -1 gpurun_out/r2c_bench_$1$2.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_bench_$1$2.json").read().strip().splitlines()[-1])
    print("rank_v2=$1 fused_update=$2:", round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"],
          {k:(round(v["ms"],4), round(v["frac"],3) if v.get("frac") else None) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"], "loss", d["loss"])
except Exception as e:
    print("no result", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rank_fused2" -s 3 -c 1 -f -o gpurun_out/r2_prof_rank2 \
    python scripts/profile_step.py --precision f16x3 --steps 4 > gpurun_out/r2_prof_rank2.log 2>&1; echo "ncu rc=$?"
