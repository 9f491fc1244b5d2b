#!/bin/bash
# exposed all-reduce time of the 8-GPU step under different NCCL algorithm choices
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 150 --warmup 5 --no-cpu-baseline > gpurun_out/nccl_$tag.json 2> gpurun_out/nccl_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/nccl_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"]), "ms/step %.3f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "allreduce %.4f"%d["kernels"]["allreduce"]["ms"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag failed", e)
PY
}
run default NCCL_DEBUG=WARN
# per-function syntax: a bare NCCL_ALGO=NVLS also binds torch's own broadcast, which has no NVLS variant
run nvls NCCL_ALGO=allreduce:NVLS
run ring_simple NCCL_ALGO=allreduce:Ring NCCL_PROTO=Simple
