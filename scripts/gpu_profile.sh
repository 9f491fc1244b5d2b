#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for p in tf32x3 bf16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$p.csv \
      python scripts/profile_step.py --precision $p --steps 4 > gpurun_out/launches_$p.log 2>&1; echo "launch list $p rc=$?"
done
# full captures of the dominant kernels (tf32x3 = default bench mode), step 3 of 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -o gpurun_out/prof_gemm_tf32x3 \
    python scripts/profile_step.py --precision tf32x3 --steps 4 > gpurun_out/prof_gemm_tf32x3.log 2>&1; echo "full gemm x3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -o gpurun_out/prof_gemm_bf16 \
    python scripts/profile_step.py --precision bf16 --steps 4 > gpurun_out/prof_gemm_bf16.log 2>&1; echo "full gemm bf16 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rank_|gather_rows|sgd_update" -s 10 -c 5 -o gpurun_out/prof_stream_tf32x3 \
    python scripts/profile_step.py --precision tf32x3 --steps 4 > gpurun_out/prof_stream_tf32x3.log 2>&1; echo "full stream rc=$?"
ls -la gpurun_out/
