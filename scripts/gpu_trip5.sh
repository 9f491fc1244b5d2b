#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 120 python scripts/diag_gemm.py tf32x3 > gpurun_out/diag_tf32x3.log 2>&1; echo "diag rc=$?"; grep -c "OK " gpurun_out/diag_tf32x3.log; grep -A8 BAD gpurun_out/diag_tf32x3.log | head -40
timeout 100 python scripts/diag_x3.py 2>&1 | tail -9
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "^E    \+where\|^E   +" gpurun_out/pytest_gpu.log | tail -n 25
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","clocks")}); print("e2e",d["e2e"]["value"])
    print({k:(round(v["ms"],4), round(v.get("frac") or 0,3)) for k,v in d["kernels"].items()})
except Exception as e: print("no bench line", e)
PY
