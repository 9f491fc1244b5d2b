#!/bin/bash
# cta_group::2 forward: correctness (gather-fused trainer tests, full size) then A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export VV_GEMM_2CTA=1
timeout 240 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_fullsize.py -q -x --timeout 60 -k "gather_fused or bench_configuration or large_window" 2>&1 | tail -6
for v in 1 0 1 0; do
  VV_GEMM_2CTA=$v timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2cta=$v', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'], d['loss'])"
done
for v in 1 0; do
  VV_GEMM_2CTA=$v timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 2cta=$v', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['loss'])"
done
