#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gather_fused" 2>&1 | tail -15
for p in f16x3 bf16; do
  timeout 300 python bench.py --steps 100 --warmup 5 --precision $p --no-cpu-baseline --fused-gather > gpurun_out/bench_${p}_fg.json 2> gpurun_out/bench_${p}_fg.err
  tail -3 gpurun_out/bench_${p}_fg.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${p}_fg.json").read().strip().splitlines()[-1])
print("$p fused-gather", round(d["value"]), "ms/step %.3f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), {k:(round(v["ms"],4), round(v.get("frac") or 0,3)) for k,v in d["kernels"].items()}, "loss", d["loss"])
PY
done
