#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for p in f16x3 bf16; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_step.py $p > gpurun_out/san_${tool}_$p.log 2>&1
    echo "$tool $p rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_$p.log | tail -1) $(grep -c 'steps ok' gpurun_out/san_${tool}_$p.log)"
  done
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_step.py tf32x3 mat > gpurun_out/san_memcheck_tf32x3.log 2>&1; echo "memcheck tf32x3 rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/san_memcheck_tf32x3.log | tail -1)"
grep -B2 -A12 "Invalid\|Race\|hazard\|Barrier error" gpurun_out/san_*.log | head -60
