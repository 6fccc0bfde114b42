#!/bin/bash
# scripts/gpurun_retry.sh <log> <timeout> [--gpus N] -- '<command>' : retries while the pod answers busy (exit 3)
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|no box or slot\|retry in a few minutes" $LOG; then sleep 90; continue; fi
  exit $rc
done
exit 3
