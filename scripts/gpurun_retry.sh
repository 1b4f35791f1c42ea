#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3 or status=transient).
# Usage: scripts/gpurun_retry.sh <logfile> <timeout_s> <command string> [--gpus N]
LOG=$1; TMO=$2; CMD=$3; shift 3
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$TMO" -- "$CMD" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc attempt=$attempt" >> "$LOG"
