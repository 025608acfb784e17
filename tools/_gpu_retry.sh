#!/bin/bash
# usage: [GPURUN_FLAGS="--gpus 2"] tools/_gpu_retry.sh <log> <timeout> <cmd> ; retries while the pod is busy (exit code 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then echo "gpurun rc=$rc" >> $log; exit $rc; fi
  sleep 60
done
