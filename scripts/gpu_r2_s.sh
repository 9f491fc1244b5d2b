#!/bin/bash
# compute-sanitizer memcheck over the trainer's gather-fused path (cta_group::2 GEMMs, fused finish, rank loss) and the
# data-layer option tests, small shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_caffe_host.py -q --timeout 90 -x \
  -k "gather_fused or rand_skip" > gpurun_out/r2s_memcheck.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|Timeout" gpurun_out/r2s_memcheck.log | head -12
