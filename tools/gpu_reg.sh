#!/bin/bash
# Registration-kernel iteration: the parity tests that go through k_register, then the odometry bench (kernel table) without the extra legs.
#   LOOP=1 adds the loop_batch leg (bench default legs);  NCU=1 adds a full ncu capture of k_register
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_register_gpu.py tests/test_odom_gpu.py tests/test_loopdb_gpu.py tests/test_loop_gpu.py tests/test_cpp_host_gpu.py tests/test_coral_gpu.py \
   tests/test_envelope_gpu.py -m gpu -x -q > gpurun_out/pytest_reg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_reg.log
tail -4 gpurun_out/pytest_reg.log
if [ "${LOOP:-0}" = "1" ]; then EXTRA=""; else EXTRA="--no-extra-legs"; fi
timeout 600 python bench.py --no-cpu-baseline $EXTRA --steps 10 --warmup 3 > gpurun_out/bench_reg.json 2> gpurun_out/bench_reg.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_reg.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()}, "parity", d.get("parity"))
    if "loop_batch" in d:
        for k, v in d["loop_batch"]["batches"].items():
            print("loop", k, v["value"], v["ms_per_iter"], v["k_register_ms"], v["device_ms"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_reg.err
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_register -s 6 -c 1 -f -o gpurun_out/full_k_register \
     python bench.py --no-cpu-baseline --no-extra-legs --steps 2 --warmup 5 > gpurun_out/ncu_full_k_register.log 2>&1; echo "ncu rc=$?"
fi
