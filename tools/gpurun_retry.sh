#!/bin/bash
# retry a gpurun call while the pod answers "busy / draining" (nothing is charged for those); usage: gpurun_retry.sh TIMEOUT 'command'
T=$1; shift
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box or slot"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
