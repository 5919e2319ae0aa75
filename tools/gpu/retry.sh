#!/bin/bash
# usage: tools/gpu/retry.sh <logfile> <timeout> <command...> : re-submits while the pod answers busy (exit 3 / transient)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|status=busy" $LOG || grep -q "retry in a few minutes" $LOG; then sleep 75; continue; fi
  break
done
