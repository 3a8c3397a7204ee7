#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> '<command>'  [--gpus N]
# retries gpurun while the pod answers "busy/transient" (exit 3); prints the tail of the output
T=$1; CMD=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then
    echo "[gpu_retry] attempt $attempt: busy, retrying in 90 s"; sleep 90; continue
  fi
  break
done
tail -80 /tmp/gpurun_last.log
exit $rc
