#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for p in tf32x3 bf16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$p.csv \
      python scripts/profile_step.py --precision $p --steps 4 > gpurun_out/launches_$p.log 2>&1; echo "launch list $p rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rank_fused" -s 2 -c 1 -o gpurun_out/prof_rank_$p \
    python scripts/profile_step.py --precision $p --steps 4 > gpurun_out/prof_rank_$p.log 2>&1; echo "full rank $p rc=$?"
done
timeout 300 python -m pytest tests/test_gpu_caffe_host.py -x -q 2>&1 | tail -3
ls -la gpurun_out/
