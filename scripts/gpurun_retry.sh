#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout_s> <command...>   -- retries while the pod answers busy (rc 3 / transient)
LOG=$1; TO=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpurun_retry: rc=$rc attempts=$i" >> $LOG
