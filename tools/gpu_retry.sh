#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers busy/transient (rc 3)
log=$1; shift
for attempt in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy\|no box\|rc=3" "$log" && ! grep -q "status=ok" "$log"; then
    sleep 90
    continue
  fi
  exit $rc
done
