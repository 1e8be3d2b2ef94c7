#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = nothing charged): scripts/gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] pod busy, sleeping 90 s"; sleep 90
done
exit 3
