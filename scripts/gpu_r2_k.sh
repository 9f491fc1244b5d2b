#!/bin/bash
# cta_group::2 weight gradient: correctness (trainer, full size) then A/B; plus the weighted max-margin test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_golden.py -q -x --timeout 100 -k "weighted or layer_kernels" 2>&1 | tail -4
export VV_GEMM_2CTA_WGRAD=1
timeout 300 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_fullsize.py -q -x --timeout 100 -k "gather_fused or bench_configuration or large_window" 2>&1 | tail -6
for v in 1 0 1 0; do
  VV_GEMM_2CTA_WGRAD=$v timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wgrad2cta=$v', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'], d['loss'])"
done
for v in 1 0; do
  VV_GEMM_2CTA_WGRAD=$v timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 wgrad2cta=$v', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['loss'])"
done
