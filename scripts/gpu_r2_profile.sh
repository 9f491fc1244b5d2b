#!/bin/bash
# round-2 profile set: the launch list of bench.py itself + full captures of the GEMMs (fused split-K finish / drain, forward
# tail split), the second-generation rank-loss kernel, bf16 GEMMs.  Files land in gpurun_out/r02/ (scripts/summarize_ncu.py r02).
cd "$(dirname "$0")/.."
O=gpurun_out/r02
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $O/launches_bench.log 2>&1; echo "launch list bench rc=$?"
for p in f16x3 bf16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -f -o $O/prof_gemm_$p \
      python scripts/profile_step.py --precision $p --steps 4 > $O/prof_gemm_$p.log 2>&1; echo "full gemm $p rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rank_fused2|gather_plan" -s 5 -c 2 -f -o $O/prof_stream_f16x3 \
    python scripts/profile_step.py --precision f16x3 --steps 4 > $O/prof_stream_f16x3.log 2>&1; echo "full stream rc=$?"
ls -la $O | head -20
