#!/bin/bash
# round 2, trip A (1 GPU): whole GPU test-suite, smoke, default bench (headline + configs[2..4])
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | head -2
nproc
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2a_smoke.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2a_bench.json").read().strip().splitlines()[-1])
print("headline", round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("ring"), d["config"].get("dp_mode"))
print("  kernels", {k:(round(v["ms"],4), round(v["frac"],3) if v.get("frac") else None) for k,v in d["kernels"].items()})
print("  roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","tensor_pipe_frac")}, "clocks", d["clocks"])
for k,v in d.get("configs",{}).items():
    print(k, round(v["value"]), "ms %.4f"%v["ms_per_step"], "roofline", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"]),
          {kk:(round(vv["ms"],4), round(vv["frac"],3) if vv.get("frac") else None) for kk,vv in v.get("kernels",{}).items()})
print("cpu", d.get("cpu_baseline",{}).get("value"))
PY
