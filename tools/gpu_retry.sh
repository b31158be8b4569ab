#!/bin/bash
# usage: tools/gpu_retry.sh <gpus> <timeout> '<command>'   -- retries gpurun while the pod answers busy (rc 3), logs to gpurun_out/retry.log
G=$1; T=$2; shift 2
mkdir -p gpurun_out
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/retry.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > gpurun_out/retry.log 2>&1; fi
  rc=$?
  if ! grep -q "status=transient" gpurun_out/retry.log; then exit $rc; fi
  sleep 60
done
exit 3
