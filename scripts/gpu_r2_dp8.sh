#!/bin/bash
# round 2, 8-GPU trip: one DP_CHECK, then the scaling bench (peer-memory exchange, full default run), NCCL mode for comparison
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
nproc; nvidia-smi --query-gpu=name,power.limit --format=csv,noheader | sort | uniq -c
VV_DP_TIMEOUT_MS=5000 timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  scripts/dp_check.py f16x3 2>&1 | grep -E "DP_CHECK|iter|no-update|Error|error|Traceback" | tail -8 | tee gpurun_out/dp_check_$N.log
if ! grep -q "DP_CHECK" gpurun_out/dp_check_$N.log; then echo "no DP_CHECK verdict: aborting the trip"; exit 1; fi
show() {
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1])
    print("$1 gpus", d["n_gpus"], round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["config"].get("dp_mode"),
          {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "streams", d["config"]["sampler_note"][-40:])
    for k,v in d.get("configs",{}).items():
        print("   ", k, round(v["value"]), "ms %.4f"%v["ms_per_step"], "e2e", round(v["e2e"]["value"]), v.get("config",{}).get("dp_mode"))
except Exception as e:
    print("$1: no result", e)
PY
}
timeout -k 5 240 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/scale_r2_1.json 2> gpurun_out/scale_r2_1.err; show scale_r2_1
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/scale_r2_${N}_p2p.json 2> gpurun_out/scale_r2_${N}_p2p.err; tail -2 gpurun_out/scale_r2_${N}_p2p.err | cut -c1-300; show scale_r2_${N}_p2p
VV_DP_MODE=nccl timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 5 --no-extra-configs > gpurun_out/scale_r2_${N}_nccl.json 2> gpurun_out/scale_r2_${N}_nccl.err; show scale_r2_${N}_nccl
