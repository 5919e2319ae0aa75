#!/bin/bash
O=gpurun_out/${1:-nculm}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
ensure_model('birdnet_v30')"
FAMILY=birdnet_v30 B=512 REPS=1 timeout 300 python tools/run_once.py > $O/run_once.log 2>&1; tail -2 $O/run_once.log
FAMILY=birdnet_v30 B=512 REPS=1 timeout 900 ncu --set full --clock-control none -k regex:k_logmel -c 1 -o $O/lm python tools/run_once.py > $O/ncu.log 2>&1
timeout 300 ncu -i $O/lm.ncu-rep --page raw --csv > $O/lm_raw.csv 2> $O/lm_raw.err
timeout 300 ncu -i $O/lm.ncu-rep --page details > $O/lm_details.txt 2>&1
rm -f $O/lm.ncu-rep; ls -la $O
