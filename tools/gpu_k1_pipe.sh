#!/bin/bash
# K1 inside the odometry step (with motion compensation): per-kernel times + ncu capture of the fused filter kernel
set -u
mkdir -p gpurun_out
timeout 300 python tools/odom_profile.py 592 12 3 2>&1 | tail -1 | tee gpurun_out/odom_profile.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_filter_fused -s 9 -c 1 -f -o gpurun_out/full_k1_pipe python tools/odom_profile.py 592 12 1 > gpurun_out/ncu_k1_pipe.log 2>&1; echo "ncu rc=$?"
