#!/bin/bash
# round 2, 1-GPU trip after the cta_group::2 forward became the f16x3 default: whole GPU suite, smoke, default bench,
# launch list + full capture of the GEMMs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2g_smoke.log
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -2 gpurun_out/r2g_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print("headline", round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("ring"), "launches/step", d["gpu_launches"]/d["steps"])
print("  kernels", {k:(round(v["ms"],4), round(v["frac"],3) if v.get("frac") else None) for k,v in d["kernels"].items()})
print("  roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","tensor_pipe_frac","traffic")}, "clocks", d["clocks"])
for k,v in d.get("configs",{}).items():
    print(k, round(v["value"]), "ms %.4f"%v["ms_per_step"], "roofline", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"]),
          {kk:(round(vv["ms"],4), round(vv["frac"],3) if vv.get("frac") else None) for kk,vv in v.get("kernels",{}).items()})
print("cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"))
PY
O=gpurun_out/r02
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $O/launches_bench.log 2>&1; echo "launch list bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -f -o $O/prof_gemm_f16x3 \
    python scripts/profile_step.py --precision f16x3 --steps 4 > $O/prof_gemm_f16x3.log 2>&1; echo "full gemm f16x3 rc=$?"
