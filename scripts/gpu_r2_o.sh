#!/bin/bash
# last sanity of the round: smoke, the full-size / trainer parity tests, and two long identical runs (3000 steps) whose
# final losses must agree to the last bit (a synchronisation race in the GEMMs would show up as run-to-run noise)
cd "$(dirname "$0")/.."
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_trainer.py tests/test_gpu_dropin.py -q --timeout 120 2>&1 | tail -2
for i in 1 2; do
  timeout 200 python bench.py --steps 3000 --warmup 5 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i', round(d['value']), '%.4f'%d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'loss', repr(d['loss']), d['clocks']['sm_mhz'])"
done
for i in 1 2; do
  timeout 200 python bench.py --steps 3000 --warmup 5 --no-cpu-baseline --no-extra-configs --precision bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 run $i', round(d['value']), '%.4f'%d['ms_per_step'], 'loss', repr(d['loss']))"
done
