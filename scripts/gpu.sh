#!/bin/bash
# gpurun with retry while the pod has no free slot (exit code 3).  Usage: scripts/gpu.sh <timeout-seconds> '<command>' [extra gpurun flags]
T=$1; shift; CMD=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout $T -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
