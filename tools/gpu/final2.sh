#!/bin/bash
# round-end evidence at HEAD, one GPU: smoke, all GPU tests, profile set, every bench configuration, 32 kHz launch lists
O=${1:-final2}; mkdir -p gpurun_out/$O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$O/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/$O/smoke.log; tail -3 gpurun_out/$O/smoke.log
bash tools/gpu/tests.sh $O
bash tools/gpu/profile.sh $O
bash tools/gpu/benches.sh $O
bash tools/gpu/launch32k.sh $O
