#!/bin/bash
# usage: scratch/gpu.sh <log> <timeout-seconds> <command...>   retries while the pod answers busy
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|retry in a few minutes\|no box\|busy" $log && ! grep -q "charged=[1-9]" $log; then
    sleep 90; continue
  fi
  break
done
echo "__done__" >> $log
