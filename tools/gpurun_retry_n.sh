#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <log> <timeout-seconds> <command...>
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 150
done
exit 3
