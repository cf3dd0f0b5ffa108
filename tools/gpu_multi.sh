#!/bin/bash
# Multi-GPU evidence on N GPUs of one box: bench.py under torchrun (sequences sharded, no collective) and the sharded loop
# registration (NCCL all-gather of constraint records).  Usage: bash tools/gpu_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 600 $TR --master-port 29502 tests/tools/loop_bench.py --pairs 1024 --keyframes 128 > gpurun_out/loop_n$N.json 2> gpurun_out/loop_n$N.err; echo "loop N=$N rc=$?"
tail -c 1200 gpurun_out/loop_n$N.json; tail -3 gpurun_out/loop_n$N.err
timeout 600 python tests/tools/loop_bench.py --pairs 1024 --keyframes 128 > gpurun_out/loop_n1_1024.json 2> gpurun_out/loop_n1_1024.err; echo "loop N=1 rc=$?"
tail -c 600 gpurun_out/loop_n1_1024.json
