#!/bin/bash
# usage: retry_gpurun.sh <logfile> <timeout> <gpus> <command...>   retries while the pod answers transient/busy (rc 3)
log=$1; to=$2; gpus=$3; shift 3
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > $log 2>&1; fi
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|busy" $log && ! grep -q "status=ok" $log; then echo "attempt $i transient" >> $log.attempts; sleep 90; continue; fi
  break
done
echo finished rc=$rc >> $log
