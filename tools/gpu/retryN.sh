#!/bin/bash
# usage: tools/gpu/retryN.sh <gpus> <logfile> <timeout> <command> : like retry.sh on an N-GPU box
N=$1; shift; LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $TMO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|status=busy" $LOG || grep -q "retry in a few minutes\|right now" $LOG; then sleep 90; continue; fi
  break
done
