#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged).  Usage: scripts/gpurun_retry.sh <timeout_s> <cmd...>
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -70
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then
    echo "[retry $i] transient, sleeping 45 s"; sleep 45; continue
  fi
  break
done
