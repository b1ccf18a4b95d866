#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   (retries while the pod answers busy)
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log" || grep -q "rc=3" "$log"; then sleep 120; continue; fi
  break
done
