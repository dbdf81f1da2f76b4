#!/bin/bash
# usage: [GPUS=n] tools/gpurun_retry.sh <timeout_s> <out_file> <command...>   -- retries while the pod answers "transient" (nothing charged)
T=$1; OUT=$2; shift 2
G=""; if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > $OUT 2>&1
  if grep -q "status=transient" $OUT; then sleep 90; continue; fi
  break
done
