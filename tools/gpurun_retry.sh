#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <max tries> <gpurun args...>   -- retries while the pod answers "transient" (busy, nothing charged)
LOG=$1; TRIES=$2; shift 2
for i in $(seq 1 $TRIES); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then exit 0; fi
  echo "[retry $i] transient, sleeping" >> "$LOG.retries"
  sleep 150
done
exit 3
