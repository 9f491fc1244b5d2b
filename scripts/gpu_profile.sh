#!/bin/bash
# round-1 profile set for the default path (f16x3, gather fused) + bf16: launch lists and full captures
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for p in f16x3 bf16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$p.csv \
      python scripts/profile_step.py --precision $p --steps 4 > gpurun_out/launches_$p.log 2>&1; echo "launch list $p rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -o gpurun_out/prof_gemm_$p \
      python scripts/profile_step.py --precision $p --steps 4 > gpurun_out/prof_gemm_$p.log 2>&1; echo "full gemm $p rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rank_fused|gather_plan|sgd_update" -s 8 -c 4 -o gpurun_out/prof_stream_f16x3 \
    python scripts/profile_step.py --precision f16x3 --steps 4 > gpurun_out/prof_stream_f16x3.log 2>&1; echo "full stream rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_rows" -s 2 -c 1 -o gpurun_out/prof_gather_f16x3 \
    python scripts/profile_step.py --precision f16x3 --materialised --steps 4 > gpurun_out/prof_gather_f16x3.log 2>&1; echo "full gather rc=$?"
ls -la gpurun_out/ | head -30
