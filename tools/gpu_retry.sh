#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <out_file> <gpus> <command...>   -- retries while the pod answers "transient"
T=$1; OUT=$2; G=$3; shift 3
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > $OUT 2>&1; fi
  if grep -q "status=transient" $OUT || grep -q "rc=3" $OUT; then sleep 90; continue; fi
  break
done
