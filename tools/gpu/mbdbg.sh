#!/bin/bash
O=gpurun_out/${1:-mbdbg}; mkdir -p $O
for d in 0 64 0 64; do
  echo "== BN_MB_DEBUG=$d" | tee -a $O/phase_dbg.log
  BN_MB_DEBUG=$d BN_MB_PROFILE=1 timeout 300 python tools/mb_phase_profile.py 2>&1 | grep "s5b1\|s4b1" | tee -a $O/phase_dbg.log
done
