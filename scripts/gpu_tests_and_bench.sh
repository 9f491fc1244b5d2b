#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for p in tf32x3 bf16; do
  timeout 300 python bench.py --steps 100 --warmup 5 --precision $p > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$p.json").read().strip().splitlines()[-1])
print("$p", d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e", d["e2e"]["value"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
print("   roofline", d["roofline"], "cpu", d.get("cpu_baseline"))
PY
done
