#!/bin/bash
O=gpurun_out/${1:-mb}; mkdir -p $O
timeout 600 python tools/mbconv_check.py --batch 8 --times-batch 256 > $O/mbconv_check.log 2>&1; echo "exit $?" >> $O/mbconv_check.log
tail -32 $O/mbconv_check.log
BN_MB_PROFILE=1 timeout 300 python tools/mb_phase_profile.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" | tail -12 | tee $O/phase.log
