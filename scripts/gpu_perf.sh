#!/bin/bash
# quick perf sweep of the GEMM knobs (no parity checks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pf in 0 4 8 16; do for p in tf32x3 bf16; do
  VV_GEMM_PREFETCH=$pf timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --precision $p > gpurun_out/b.log 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/b.log").read().strip().splitlines()[-1])
print("pf=$pf $p ms/step %.3f fwd %.4f wgrad %.4f" % (d["ms_per_step"], d["kernels"]["fc7_forward"]["ms"], d["kernels"]["wgrad"]["ms"]))
PY
done; done
