#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trainer.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -q -x --timeout 120 -k "ip_ or gemm or trainer or step or bench_configuration or large_window or curve or golden or fused" 2>&1 | grep -E "Error|error|FAILED|failed|assert|^E " | head -20
VV_GEMM_2CTA=0 VV_GEMM_2CTA_WGRAD=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-configs --precision bf16 2>&1 | tail -5 | cut -c1-400
