#!/bin/bash
# memcheck + synccheck + racecheck of the default step (f16x3, gather fused), with the register-resident and the ring rank kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ring in 0 2; do
  for tool in memcheck synccheck racecheck; do
    VV_RANK_RING=$ring timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_step.py f16x3 > gpurun_out/san_${tool}_ring$ring.log 2>&1
    echo "ring=$ring $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_ring$ring.log | tail -1) $(grep -c 'steps ok' gpurun_out/san_${tool}_ring$ring.log)"
  done
done
grep -B2 -A12 "Invalid\|Barrier error" gpurun_out/san_*ring*.log | head -40
grep -h "Race reported\|hazard" gpurun_out/san_racecheck_ring*.log | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -12
