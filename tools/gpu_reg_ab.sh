#!/bin/bash
# A/B of k_register builds: every devlibs/libtbv_*.so in turn takes the library's place for one short odometry bench (kernel table on stdout,
# the development timers — when the variant was built with -DTBV_DEV_TIMERS — on stderr).  Scratch tool for GPU visits; devlibs/ is not tracked.
set -u
mkdir -p gpurun_out
cp tbv_slam_public_b200/libtbv_b200.so /tmp/libtbv_b200.orig.so
for v in devlibs/libtbv_*.so; do
  n=$(basename $v .so)
  cp $v tbv_slam_public_b200/libtbv_b200.so
  timeout 300 python bench.py --no-cpu-baseline --no-extra-legs --steps 6 --warmup 3 > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err; echo "$n rc=$?"
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", d["value"], "ms/step", d["ms_per_step"], "k_register", d["kernels"]["k_register"]["ms_per_step"])
except Exception as e:
    print(n, "parse failed", e)
PY
  grep "cycles/problem\|distribution" gpurun_out/ab_$n.err | tail -4
  for q in ${AB_SEQS:-}; do
    timeout 300 python bench.py --no-cpu-baseline --no-extra-legs --steps 6 --warmup 3 --seqs $q > gpurun_out/ab_${n}_$q.json 2> gpurun_out/ab_${n}_$q.err
    echo "  seqs=$q: $(python -c "import json;d=json.loads(open('gpurun_out/ab_${n}_$q.json').read().strip().splitlines()[-1]);print('k_register',d['kernels']['k_register']['ms_per_step'],'ms/step',d['ms_per_step'])" 2>&1 | tail -1)"
    grep "cycles/problem" gpurun_out/ab_${n}_$q.err | tail -1
  done
done
cp /tmp/libtbv_b200.orig.so tbv_slam_public_b200/libtbv_b200.so
