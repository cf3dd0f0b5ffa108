#!/bin/bash
# Host topology (NUMA / PCIe placement of the GPUs) + K1 baseline numbers and a source-level ncu capture.
set -u
mkdir -p gpurun_out
{
  echo "== nproc"; nproc; echo "== affinity"; taskset -p $$; grep -i "allowed" /proc/self/status
  echo "== lscpu"; lscpu | head -40
  echo "== numa nodes"; ls /sys/devices/system/node/ 2>&1; for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== numactl"; numactl -H 2>&1 | head -20
  echo "== topo"; nvidia-smi topo -m 2>&1
  echo "== gpu pci numa"; for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do b=$(echo $d | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "$d numa=$(cat /sys/bus/pci/devices/$b/numa_node 2>&1) local_cpus=$(cat /sys/bus/pci/devices/$b/local_cpulist 2>&1)"; done
  echo "== cgroup cpuset"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>&1
  echo "== mem"; free -g
} > gpurun_out/topology.txt 2>&1
for kind in radar uniform equal; do timeout 300 python tools/k1_bench.py 592 10 $kind 2>&1 | tail -1; done | tee gpurun_out/k1_bench_before.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_kstrongest -s 3 -c 1 -f -o gpurun_out/full_k1 python tools/k1_bench.py 592 2 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_register -s 9 -c 1 -f -o gpurun_out/full_k_register python tools/odom_profile.py 592 12 1 > gpurun_out/ncu_reg.log 2>&1; echo "ncu reg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cells_fused -s 9 -c 1 -f -o gpurun_out/full_cells python tools/odom_profile.py 592 12 1 > gpurun_out/ncu_cells.log 2>&1; echo "ncu cells rc=$?"
ls -la gpurun_out
