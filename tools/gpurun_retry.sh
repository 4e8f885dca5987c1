#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "transient" (busy), up to ~60 minutes
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 60; else exit 0; fi
done
exit 3
