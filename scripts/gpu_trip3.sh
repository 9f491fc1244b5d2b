#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches')}); print('e2e',d['e2e']['value']); print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['kind'],d['cpu_baseline']['cores'])
print({k:(round(v['ms'],4), round(v.get('frac') or 0,3)) for k,v in d['kernels'].items()})
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 600 gpurun_out/bench_ref.log
