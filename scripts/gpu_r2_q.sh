#!/bin/bash
# closing run of round 2: whole GPU suite, smoke, the driver's bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/r2q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2q_smoke.log
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -2 gpurun_out/r2q_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
print("headline", round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"]/d["steps"], "cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"])
print("  kernels", {k:(round(v["ms"],4), round(v["frac"],3) if v.get("frac") else None) for k,v in d["kernels"].items()}, "tensor_pipe_frac", round(d["roofline"]["tensor_pipe_frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for k,v in d.get("configs",{}).items():
    print(k, round(v["value"]), "ms %.4f"%v["ms_per_step"], "roofline", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"]))
PY
