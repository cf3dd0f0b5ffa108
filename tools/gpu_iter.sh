#!/bin/bash
# Quick GPU iteration: parity tests, then the per-kernel profile of the odometry pipeline for a few settings.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for c in 3 4; do
  for n in 512 592; do
    echo "== TBV_REG_CTAS=$c seqs=$n"
    TBV_REG_CTAS=$c timeout 300 python tools/odom_profile.py $n 12 3 2>&1 | tail -1
  done
done | tee gpurun_out/odom_profile.log
timeout 600 python tools/loop_bench.py 2>&1 | tail -2 | tee gpurun_out/loop_bench_1gpu.json
