#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers "transient"/busy)
T=$1; shift
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box or slot"; then
    echo "[retry $attempt] pod busy, sleeping 90 s" >&2
    sleep 90
    continue
  fi
  echo "$out"
  exit 0
done
echo "$out"
exit 3
