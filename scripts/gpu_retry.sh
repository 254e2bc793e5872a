#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> '<command>'  -- retries gpurun while the pod answers busy (nothing is charged)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gpu_retry: still busy after 30 attempts"; exit 3
