#!/bin/bash
# usage: gpu_retry.sh <timeout> <command...>   retries while the pod is busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gave up"; exit 3
