#!/bin/bash
# round 2, final 1-GPU trip: whole GPU test-suite, smoke, the default bench (driver's command), the reference arm, profiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print("headline", round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("ring"), "launches/step", d["gpu_launches"]/d["steps"])
print("  kernels", {k:(round(v["ms"],4), round(v["frac"],3) if v.get("frac") else None) for k,v in d["kernels"].items()})
print("  roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","tensor_pipe_frac","traffic")}, "clocks", d["clocks"])
for k,v in d.get("configs",{}).items():
    print(k, round(v["value"]), "ms %.4f"%v["ms_per_step"], "roofline", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"]),
          {kk:(round(vv["ms"],4), round(vv["frac"],3) if vv.get("frac") else None) for kk,vv in v.get("kernels",{}).items()})
print("cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"))
r=json.loads(open("gpurun_out/r2f_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["config"]["global_batch"], r["cpu_baseline"]["cores"])
PY
bash scripts/gpu_r2_profile.sh 2>&1 | tail -12
