#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> [gpurun args...] -- retries while the pod answers "transient" (exit 3), at most 15 times
log=$1; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
