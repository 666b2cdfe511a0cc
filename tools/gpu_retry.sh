#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <gpurun args...>
# Retries only while the pod has no box / slot free (gpurun exit code 3 with nothing charged, or a
# "transient" verdict); any real verdict -- ok or fail -- ends the loop.
log=$1; shift
for attempt in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy\|status=refused" "$log"; then
    sleep 60
    continue
  fi
  exit $rc
done
