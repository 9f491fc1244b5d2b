#!/bin/bash
# MMA issuer on the uniform datapath: GEMM correctness, then timings (f16x3; bf16 with and without the cta_group::2 forms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trainer.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -q -x --timeout 120 -k "ip_ or gemm or trainer or step or bench_configuration or large_window or curve or golden or fused" 2>&1 | tail -5
for i in 1 2; do
timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f16x3', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'], d['loss'])"
done
for v in "0 0" "1 1" "1 0" "0 1"; do
  set -- $v
  VV_GEMM_2CTA=$1 VV_GEMM_2CTA_WGRAD=$2 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 2cta fwd=$1 wgrad=$2', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['loss'])"
done
