#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pf in 0 1 2 4; do
  VV_GATHER_PREFETCH=$pf timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b.log 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/b.log").read().strip().splitlines()[-1])
print("pf=$pf ms/step %.3f fwd %.4f wgrad %.4f" % (d["ms_per_step"], d["kernels"]["fc7_forward"]["ms"], d["kernels"]["wgrad"]["ms"]))
PY
done
