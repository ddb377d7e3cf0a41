#!/bin/bash
# usage: gpu_retry.sh <log> <timeout_s> <command...>  -- retries a gpurun call while the pod answers "transient" (nothing charged)
log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 60; else break; fi
done
