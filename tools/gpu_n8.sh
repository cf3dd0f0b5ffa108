#!/bin/bash
# bench.py on 4 GPUs of one box (sequences sharded; loop_batch / mulran legs with the pipelined sharded registration)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 bench.py --no-cpu-baseline --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench N=8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("h2d_gbs_per_gpu"))
for k, v in d["loop_batch"]["batches"].items():
    print("  loop", k, v["value"], v["ms_per_iter"], v["device_ms"], v["allgather_us"])
print("  mulran", d["mulran"]["odometry_alone"], d["mulran"]["mixed"])
PY
tail -n 3 gpurun_out/bench_n8.err
