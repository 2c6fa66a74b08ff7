#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (nothing is charged for those).
# usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -120
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 60; continue; fi
  break
done
