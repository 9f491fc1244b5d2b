#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_trainer.py -q -x -k "gather_fused" > gpurun_out/pytest_gf.log 2>&1; echo "pytest rc=$?"; grep -v "^E    \+where\|^E   +" gpurun_out/pytest_gf.log | tail -n 30
for p in tf32x3 bf16; do
timeout 600 python bench.py --steps 200 --warmup 5 --precision $p --no-cpu-baseline > gpurun_out/bench_$p.log 2>gpurun_out/bench_$p.err; echo "bench $p rc=$?"; tail -2 gpurun_out/bench_$p.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$p.log").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks")}); print("e2e",d["e2e"]["value"])
    print({k:(round(v["ms"],4), round(v.get("frac") or 0,3)) for k,v in d["kernels"].items()})
except Exception as e: print("no bench line", e)
PY
done
