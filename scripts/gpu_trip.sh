#!/bin/bash
# One GPU-box visit: staged so that a hang in a late stage still leaves the earlier logs.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for p in tf32 tf32x3; do
  echo "== diag $p" ; timeout 120 python scripts/diag_gemm.py $p > gpurun_out/diag_$p.log 2>&1; echo "rc=$?"
done
grep -c OK gpurun_out/diag_tf32.log gpurun_out/diag_tf32x3.log; grep BAD gpurun_out/diag_tf32*.log | head
echo "== pytest all gpu"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -n 60 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/bench.log
timeout 600 python bench.py --steps 100 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_bf16.log
