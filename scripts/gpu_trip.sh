#!/bin/bash
# One GPU-box visit: staged so that a hang in a late stage still leaves the earlier logs.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== diag simt" ; timeout 300 python scripts/diag_gemm.py fp32_simt > gpurun_out/diag_simt.log 2>&1; echo "rc=$?"
for p in bf16 tf32 tf32x3; do
  echo "== diag $p" ; timeout 120 python scripts/diag_gemm.py $p > gpurun_out/diag_$p.log 2>&1; echo "rc=$?"
done
tail -n 30 gpurun_out/diag_bf16.log
echo "== pytest non-GEMM"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "bank or gather or rank_loss or sgd or standalone or known_answers or errors" > gpurun_out/pytest_stream.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/pytest_stream.log
echo "== pytest all gpu"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench.log
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench_bf16.log
