#!/bin/bash
# Pipelined sharded loop registration on 2 GPUs: the comm / C++ host tests, then bench.py at N=1 and N=2 (loop_batch + mulran legs).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_cpp_host_gpu.py tests/test_loopdb_gpu.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pipe.log
tail -15 gpurun_out/pytest_pipe.log
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_p1.json 2> gpurun_out/bench_p1.err; echo "bench N=1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 bench.py --no-cpu-baseline --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_p2.json 2> gpurun_out/bench_p2.err; echo "bench N=2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_p1.json", "gpurun_out/bench_p2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
        for k, v in d["loop_batch"]["batches"].items():
            print("  loop", k, v["value"], v["ms_per_iter"], v["k_register_ms"], v["device_ms"], v["allgather_us"])
        print("  mulran", d["mulran"]["odometry_alone"], d["mulran"]["mixed"])
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -n 3 gpurun_out/bench_p1.err; tail -n 3 gpurun_out/bench_p2.err
