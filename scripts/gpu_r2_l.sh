#!/bin/bash
# bf16 forward: cta_group::2 vs multicast kernel under ncu (stall picture of the pair form)
cd "$(dirname "$0")/.."
O=gpurun_out/r02
mkdir -p $O
for v in 1 0; do
VV_GEMM_2CTA=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -f -o $O/prof_fwd_bf16_2cta$v \
    python scripts/profile_step.py --precision bf16 --steps 4 > $O/prof_fwd_bf16_2cta$v.log 2>&1; echo "rc=$?"
done
ls -la $O | tail -5
