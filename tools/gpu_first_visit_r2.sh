#!/bin/bash
# First GPU visit of the NEXT round (about 3 GPU-minutes): everything that was written after round 1's GPU budget was spent.
#   1. the whole GPU suite (new: offline odometry reader, cloud-interface fuser, batched / sharded loop-closure search);
#   2. the thread-block-cluster PCG (opt-in kernel pgo_pcg_cluster): parity against scipy + us per CG iteration, next to the one-CTA kernel;
#   3. the offline flow end to end on a synthetic drive.
# Results land in gpurun_out/; copy what should be judged into profiles/r2a_*.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python tests/tools/pgo_cluster_check.py > gpurun_out/pgo_one_cta.json 2> gpurun_out/pgo_one_cta.err; echo "pgo one-CTA rc=$?"
TBV_PGO_CLUSTER=1 timeout 120 python tests/tools/pgo_cluster_check.py > gpurun_out/pgo_cluster.json 2> gpurun_out/pgo_cluster.err; echo "pgo cluster rc=$?"
tail -c 1500 gpurun_out/pgo_cluster.json; tail -3 gpurun_out/pgo_cluster.err
TBV_PGO_CHAIN=1 timeout 120 python tests/tools/pgo_cluster_check.py > gpurun_out/pgo_chain.json 2> gpurun_out/pgo_chain.err; echo "pgo chain rc=$?"
tail -c 1500 gpurun_out/pgo_chain.json; tail -3 gpurun_out/pgo_chain.err
timeout 300 python tools/slam_offline.py --out gpurun_out/slam_offline --frames 300 > gpurun_out/slam_offline.json 2> gpurun_out/slam_offline.err; echo "slam rc=$?"
tail -c 1500 gpurun_out/slam_offline.json; tail -3 gpurun_out/slam_offline.err
rm -rf gpurun_out/slam_offline/simple_graph.tbvg gpurun_out/slam_offline/optimised_graph.tbvg   # large: clouds of every keyframe
