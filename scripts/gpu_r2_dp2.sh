#!/bin/bash
# round 2, lean 2-GPU trip after a kernel change: the DP tests, one DP_CHECK per mode, one scaling point
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
[ "$2" = nopytest ] || timeout -k 5 400 python -m pytest tests/test_gpu_dp.py -q -x --timeout 150 2>&1 | tail -4
for mode in p2p nccl; do
  VV_DP_MODE=$mode VV_DP_TIMEOUT_MS=5000 timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/dp_check.py f16x3 2>&1 | grep -E "DP_CHECK|Error|error|Traceback" | tail -4 | tee -a gpurun_out/dp_check_${N}_b.log
done
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 --no-extra-configs > gpurun_out/scale_r2b_$N.json 2> gpurun_out/scale_r2b_$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_r2b_$N.json").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], round(d["value"]), "ms/step %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["config"].get("dp_mode"),
      {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"], d.get("rank_skew"))
PY
