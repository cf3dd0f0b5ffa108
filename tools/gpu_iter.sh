#!/bin/bash
# Quick GPU iteration: parity tests, per-kernel profile of the odometry pipeline, ncu full captures of the kernels named in $NCU_KERNELS.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for n in ${SEQS:-512 592}; do
  echo "== seqs=$n"
  timeout 300 python tools/odom_profile.py $n 12 3 2>&1 | tail -1
done | tee gpurun_out/odom_profile.log
for k in ${NCU_KERNELS:-}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 9 -c 1 -f -o gpurun_out/full_$k \
     python tools/odom_profile.py 512 12 1 > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
done
