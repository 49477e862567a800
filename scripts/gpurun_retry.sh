#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" (exit 3, nothing charged)
log=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> "$log"; exit $rc; fi
  sleep 45
done
echo "gave up" >> "$log"
