#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for p in bf16 tf32 tf32x3; do timeout 60 python scripts/diag_gemm.py $p > gpurun_out/diag_$p.log 2>&1; echo "$p rc=$? ok=$(grep -c 'OK ' gpurun_out/diag_$p.log) bad=$(grep -c 'BAD' gpurun_out/diag_$p.log)"; grep -A6 BAD gpurun_out/diag_$p.log | head -20; tail -2 gpurun_out/diag_$p.log | grep -i error; done
