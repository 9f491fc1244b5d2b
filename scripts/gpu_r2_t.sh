#!/bin/bash
# refresh the bf16 GEMM capture after the MMA-issuer change
cd "$(dirname "$0")/.."
O=gpurun_out/r02
mkdir -p $O
timeout 120 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -f -o $O/prof_gemm_bf16 \
    python scripts/profile_step.py --precision bf16 --steps 4 > $O/prof_gemm_bf16.log 2>&1; echo "rc=$?"
