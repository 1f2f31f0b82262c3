#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged).
#   tools/gpurun_retry.sh <log file> <gpurun args...>
log=$1; shift
for i in $(seq 1 200); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 15
done
exit 3
