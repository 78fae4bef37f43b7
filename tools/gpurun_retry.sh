#!/bin/bash
# retry gpurun until it gets a box (exit 3 / "transient" = nothing charged); usage: gpurun_retry.sh LOG TIMEOUT 'command'
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|nothing was charged" $log; then sleep 90; continue; fi
  break
done
echo "[retry] finished after $i attempt(s)" >> $log
