#!/bin/bash
# one full ncu capture (source-level) of k_register inside the odometry bench
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_register -s 12 -c 1 -f -o gpurun_out/full_k_register \
   python bench.py --no-cpu-baseline --no-extra-legs --steps 12 --warmup 3 > gpurun_out/ncu_full_k_register.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full_k_register.log
