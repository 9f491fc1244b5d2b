#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python scripts/diag_gemm.py f16x3 > gpurun_out/diag_f16x3.log 2>&1; echo "diag rc=$? ok=$(grep -c 'OK ' gpurun_out/diag_f16x3.log) bad=$(grep -c 'BAD' gpurun_out/diag_f16x3.log)"; grep -B1 -A5 "BAD\|failed\|Error" gpurun_out/diag_f16x3.log | head -40
timeout 900 python -m pytest tests -m gpu -x -q -k "f16x3" 2>&1 | tail -15
for p in f16x3; do
  timeout 300 python bench.py --steps 100 --warmup 5 --precision $p --no-cpu-baseline > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err
  tail -3 gpurun_out/bench_$p.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$p.json").read().strip().splitlines()[-1])
print("$p", d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
done
