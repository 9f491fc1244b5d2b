#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py f16x3 2>&1 | grep -E "DP_CHECK|Error|error" | head -5
timeout 600 python -m pytest tests/test_gpu_dp.py -q 2>&1 | tail -3
for g in 1 $N; do
  if [ $g = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$g.json 2> gpurun_out/scale_$g.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $g --steps 200 --warmup 5 > gpurun_out/scale_$g.json 2> gpurun_out/scale_$g.err; fi
  tail -2 gpurun_out/scale_$g.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_$g.json").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], round(d["value"]), "ms/step %.3f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["clocks"])
PY
done
