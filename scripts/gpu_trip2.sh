#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_caffe_host.py -q -x > gpurun_out/pytest_host.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/pytest_host.log
