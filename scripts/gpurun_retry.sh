#!/bin/bash
# usage: [GPURUN_GPUS=N] scripts/gpurun_retry.sh <log> <timeout_s> '<command>'   -- retries while the pod answers busy (exit 3 / transient)
log=$1; to=$2; cmd=$3
extra=""
if [ -n "${GPURUN_GPUS:-}" ]; then extra="--gpus $GPURUN_GPUS"; fi
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun $extra --timeout "$to" -- "$cmd" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 75; continue; fi
  break
done
