#!/bin/bash
# compute-sanitizer memcheck over the kernels added at the end of round 2 (small shapes only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -q --timeout 160 -x \
  -k "(rank_loss_wide and (5-3-33 or 9-9-40)) or retrieval_stats_video or weighted_max_margin" > gpurun_out/r2r_memcheck.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2r_memcheck.log | head -12
