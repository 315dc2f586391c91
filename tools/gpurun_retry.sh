#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout-seconds> <command...>   (retries while the pod answers busy: exit code 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
