#!/bin/bash
O=gpurun_out/${1:-sanitize}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
for f in ('birdnet_v24','birdnet_v30'): ensure_model(f)"
B=4 REPS=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_once.py > $O/memcheck_v24.log 2>&1; echo "memcheck v24 exit $?"; grep -E "ERROR SUMMARY|Invalid|Error" $O/memcheck_v24.log | head -5
FAMILY=birdnet_v30 B=4 REPS=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_once.py > $O/memcheck_v30.log 2>&1; echo "memcheck v30 exit $?"; grep -E "ERROR SUMMARY|Invalid|Error" $O/memcheck_v30.log | head -5
B=300 REPS=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_once.py > $O/memcheck_v24_b300.log 2>&1; echo "memcheck v24 B=300 exit $?"; grep -E "ERROR SUMMARY|Invalid|Error" $O/memcheck_v24_b300.log | head -5
