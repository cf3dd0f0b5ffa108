#!/bin/bash
# K1 iteration: filter / golden / odometry parity, then the K1 micro-benchmark on radar and stress inputs, optional ncu capture.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_filter_gpu.py tests/test_golden_gpu.py tests/test_odom_gpu.py tests/test_cpp_host_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_k1.log
for kind in radar uniform equal; do timeout 300 python tools/k1_bench.py 592 10 $kind 2>&1 | tail -1; done | tee gpurun_out/k1_bench.jsonl
timeout 300 python tools/odom_profile.py 592 12 3 2>&1 | tail -1 | tee gpurun_out/odom_profile.json
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_filter_fused -s 3 -c 1 -f -o gpurun_out/full_k1 python tools/k1_bench.py 592 2 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
fi
