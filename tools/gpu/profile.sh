#!/bin/bash
# round-2 profiling evidence at HEAD (1 GPU): live stage times, ncu launch list of the bench command, ncu --set full of one batch
O=gpurun_out/${1:-r02prof}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
ensure_model('birdnet_v24')"
PROFILE=1 timeout 300 python tools/run_once.py > $O/stage_times_b256.txt 2>&1; tail -3 $O/stage_times_b256.txt
BN_MB_PROFILE=1 timeout 300 python tools/mb_phase_profile.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" > $O/mbconv_phase_cycles.txt
timeout 300 python tools/fe_phase_profile.py 2>&1 | grep -v "Exception\|Traceback\|File\|Attribute" > $O/fe_phase_cycles.txt; cat $O/fe_phase_cycles.txt
# the command must exit 0 without ncu first
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ingest > $O/bench_short.json 2> $O/bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ingest > $O/ncu_launches.log 2>&1
REPS=1 timeout 300 python tools/run_once.py > $O/run_once.log 2>&1 && \
REPS=1 timeout 1500 ncu --set full --clock-control none -o $O/full python tools/run_once.py > $O/ncu_full.log 2>&1
timeout 600 ncu -i $O/full.ncu-rep --page raw --csv > $O/ncu_full_raw.csv 2> $O/ncu_full_raw.err
ls -la $O; rm -f $O/full.ncu-rep
tail -2 $O/ncu_full.log
