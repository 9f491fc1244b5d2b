#!/bin/bash
# wide rank-loss kernel: occupancy / prefetch variants alone, then DRAM bytes of two of them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for occ in 2 3 4; do for pf in 0 1; do
  VV_RANK_WIDE_CTAS=$occ VV_RANK_WIDE_PREFETCH=$pf timeout 100 python scripts/bench_rank_wide.py 2>&1 | tail -1
done; done
for occ in 2 4; do
VV_RANK_WIDE_CTAS=$occ timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:rank_wide -s 3 -c 1 \
   python scripts/bench_rank_wide.py --iters 2 2>&1 | grep -E "rank_wide|dram__|gpu__time|hit_rate" | cut -c1-150
done
