#!/bin/bash
# widening tests of the end of round 2: data-layer options, video-level retrieval + CSV, weighted loss, eval fixtures
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_gpu_golden.py tests/test_gpu_caffe_host.py -q --timeout 120 -k "retrieval or eval or rand_skip or weighted or records or test_net" 2>&1 | tail -15
