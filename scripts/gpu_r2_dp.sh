#!/bin/bash
# round 2, multi-GPU trip: DP_CHECK (peer-memory exchange and NCCL) on N ranks, then the scaling bench in both modes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
STEPS=${2:-200}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12
nproc
: > gpurun_out/dp_check_$N.log
for mode in p2p nccl; do
  for prec in f16x3 tf32x3 bf16; do
    VV_DP_MODE=$mode VV_DP_TIMEOUT_MS=5000 timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/dp_check.py $prec 2>&1 | grep -E "DP_CHECK|iter|no-update|Error|error|Traceback" | tail -8 | tee -a gpurun_out/dp_check_$N.log
    # fail fast: a first run without a verdict means something structural is wrong -- stop spending GPU time
    if ! grep -q "DP_CHECK" gpurun_out/dp_check_$N.log; then echo "no DP_CHECK verdict from the first run: aborting the trip"; exit 1; fi
  done
done
run_bench() {  # name, gpus, extra env / args
  local name=$1 g=$2; shift 2
  if [ $g = 1 ]; then timeout -k 5 240 python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline --no-extra-configs "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $g --steps $STEPS --warmup 5 --no-extra-configs "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  tail -2 gpurun_out/$name.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    print("$name gpus", d["n_gpus"], round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["config"].get("dp_mode"),
          {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$name: no result", e)
PY
}
run_bench scale_r2_1 1
run_bench scale_r2_${N}_p2p $N
VV_DP_MODE=nccl run_bench scale_r2_${N}_nccl $N
run_bench scale_r2_${N}_p2p_bf16 $N --precision bf16
