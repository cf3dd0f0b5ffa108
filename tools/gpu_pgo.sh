#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_loop_gpu.py -m gpu -x -q -k "pgo" 2>&1 | tail -15 | tee gpurun_out/pytest_pgo.log
timeout 300 python tests/tools/pgo_cluster_check.py > gpurun_out/pgo_cr.json 2> gpurun_out/pgo_cr.err; echo "check rc=$?"; tail -c 600 gpurun_out/pgo_cr.json; tail -3 gpurun_out/pgo_cr.err
timeout 600 python tools/pgo_bench.py > gpurun_out/pgo_bench.json 2> gpurun_out/pgo_bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/pgo_bench.json
