#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> <command...>   -- retries while the pod answers "busy" (rc 3)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 60
done
