#!/bin/bash
# wgrad gather producers with one address per row: correctness, then bf16 / f16x3 timing
cd "$(dirname "$0")/.."
timeout 60 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_fullsize.py -q -x --timeout 50 -k "gather_fused or bench_configuration" 2>&1 | tail -2
timeout 40 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['loss'])"
timeout 40 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('f16x3', round(d['value']), '%.4f'%d['ms_per_step'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['loss'])"
