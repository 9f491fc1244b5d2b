#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rank_fused2" -s 3 -c 1 -f -o gpurun_out/r2_prof_rank2 \
    python scripts/profile_step.py --precision f16x3 --steps 4 > gpurun_out/r2_prof_rank2.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2_prof_rank2.log
ls -la gpurun_out/r2_prof_rank2.ncu-rep
