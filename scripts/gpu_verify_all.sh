#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"]), "ms/step %.3f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), {k:(round(v["ms"],4), round(v.get("frac") or 0,3)) for k,v in d["kernels"].items()})
print({k:v for k,v in d["roofline"].items() if k not in ("note",)})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["clocks"], d["gpu_launches"], d["config"]["gather"])
PY
