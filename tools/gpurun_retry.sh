#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <out_file> <command...>   -- retries while the pod answers "transient" (nothing charged)
T=$1; OUT=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1
  if grep -q "status=transient" $OUT; then sleep 90; continue; fi
  break
done
