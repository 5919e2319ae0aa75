#!/bin/bash
# cluster-multicast front-end: correctness first (short timeouts: a protocol bug traps), then timing, plain launch for comparison
O=gpurun_out/${1:-fecl}; mkdir -p $O
timeout 120 python tools/fe_check.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" | grep "fused\|rror" | head -6
timeout 120 python tools/fe_phase_profile.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" | tail -6
echo "== BN_FE_CLUSTER=0"; BN_FE_CLUSTER=0 timeout 120 python tools/fe_phase_profile.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" | tail -6 | head -3
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages or golden or full_batch or any_size" 2>&1 | tail -3
