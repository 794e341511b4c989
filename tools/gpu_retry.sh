#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-s> <command string>   (retries while the pod answers busy / transient)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  echo "$out" | tail -40; exit $rc
done
echo "gave up"; exit 3
