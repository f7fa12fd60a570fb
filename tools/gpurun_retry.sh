#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing is charged for those).  tools/gpurun_retry.sh <gpurun args>
for i in $(seq 1 14); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -45
  if echo "$out" | grep -q "status=transient\|status=busy\|no box or slot"; then sleep 120; continue; fi
  break
done
