#!/bin/bash
# usage: tools/gpu_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy / transient
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $LOG && ! grep -q "status=ok" $LOG; then sleep 45; continue; fi
  break
done
