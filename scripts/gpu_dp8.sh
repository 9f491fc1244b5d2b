#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | head -8
for g in $N; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $g --steps 200 --warmup 5 > gpurun_out/scale_$g.json 2> gpurun_out/scale_$g.err
  tail -3 gpurun_out/scale_$g.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_$g.json").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], round(d["value"]), "ms/step %.3f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["clocks"])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
