#!/bin/bash
# fused front-end: precision check, role counters, stage times, stage test
O=gpurun_out/${1:-fe}; mkdir -p $O
timeout 300 python tools/fe_check.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" | grep "fused" > $O/check.txt; cat $O/check.txt
timeout 300 python tools/fe_phase_profile.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" | tail -6 | tee $O/phase.txt
PROFILE=1 timeout 300 python tools/run_once.py > $O/stage_times.txt 2>&1; head -4 $O/stage_times.txt; grep "TOTAL\|back-to-back" $O/stage_times.txt
