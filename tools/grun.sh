#!/bin/bash
# usage: tools/grun.sh <timeout-seconds> <logfile> <command string>   — retries while the pod has no free GPU slot (exit 3)
t=$1; log=$2; shift 2
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
exit 3
