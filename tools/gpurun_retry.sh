#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while gpurun answers "no box/slot free" (exit 3)
log=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc $rc after $i tries" >> "$log"; exit $rc; fi
  sleep 120
done
echo "gave up" >> "$log"; exit 3
