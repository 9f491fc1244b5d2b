#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trainer.py -q -x --timeout 120 -k "rank or full_size or out_of_range or gather_fused_step" 2>&1 | tail -4
for v in 4 5 6 4 5 6; do
  VV_RANK2_CTAS=$v timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas=$v', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'])"
done
