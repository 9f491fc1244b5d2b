#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in "$@"; do
  timeout 300 python bench.py --steps 100 --warmup 5 --precision $p --no-cpu-baseline > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err
  tail -3 gpurun_out/bench_$p.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$p.json").read().strip().splitlines()[-1])
print("$p", round(d["value"]), "ms/step %.3f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), {k:(round(v["ms"],4), round(v.get("frac") or 0,3)) for k,v in d["kernels"].items()})
print("   roofline", {k:v for k,v in d["roofline"].items() if k not in ("note","peak_source")}, "loss", d["loss"])
PY
done
