#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers busy (rc 3)
log=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc attempt=$attempt" >> "$log"; exit $rc; fi
  sleep 90
done
echo "rc=3 gave up" >> "$log"; exit 3
