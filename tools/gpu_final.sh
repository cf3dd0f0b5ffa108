#!/bin/bash
# Final visit of a round on one GPU: all GPU tests, smoke, the bench, the ncu launch list and one full capture per top kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --no-overlap --no-cpu-baseline --no-extra-legs > gpurun_out/bench_no_overlap.json 2> gpurun_out/bench_no_overlap.err; echo "bench --no-overlap rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 5 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for k in k1_filter_fused k_register cells_fused; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/full_$k \
     python bench.py --steps 2 --warmup 5 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?"
done
