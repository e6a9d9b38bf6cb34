#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <logfile> <command...>   retries while the pod answers "busy" (exit 3)
to=$1; log=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|no box\|retry in a few minutes" $log && ! grep -q "status=ok" $log; then sleep 120; continue; fi
  break
done
exit $rc
