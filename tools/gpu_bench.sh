#!/bin/bash
# bench.py both arms on one GPU (or N with torchrun when N > 1 is given)
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  if [ "${REF:-1}" = "1" ]; then timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err; fi
else
  { nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"; for d in /sys/bus/pci/devices/*/; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa_node=$(cat $d/numa_node)"; fi; done; free -g | head -2; } > gpurun_out/topo_n$N.txt 2>&1
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
  tail -c 3500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
fi
