#!/bin/bash
# wide rank-loss kernel (R > 32): parity, then the large-window numbers with it on and off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_trainer.py -q -x --timeout 120 -k "rank_loss or large_window or bench_configuration or oracle" 2>&1 | tail -6
for v in 1 0; do
  VV_RANK_WIDE=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); v=d['configs']['large_window']
print('wide=$v', round(v['value']), '%.4f'%v['ms_per_step'], {k:(round(x['ms'],4), round(x['frac'],3) if x.get('frac') else None) for k,x in v['kernels'].items()})"
done
