#!/bin/bash
O=gpurun_out/${1:-e2edepth}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
ensure_model('birdnet_v24')"
M=models/birdnet_v24_seed0.onnx
for d in 2 3 4 5 6 8; do timeout 300 tools/_build/e2e_driver --model $M --mode ctx --depth $d --batches 60 --pinned 1 --reps 3 >> $O/depth.jsonl 2>> $O/depth.err; done
timeout 300 tools/_build/e2e_driver --model $M --mode pool --depth 5 --batches 60 --pinned 1 --reps 3 >> $O/depth.jsonl 2>> $O/depth.err
BN_TRACE_RUN=1 timeout 120 tools/_build/e2e_driver --model $M --mode ctx --depth 5 --batches 20 --pinned 1 --reps 1 > $O/trace.json 2> $O/trace.err
python - <<PY
import json
for l in open("$O/depth.jsonl"):
    d=json.loads(l); print(d["mode"], "depth", d["depth"], round(d["segments_per_s_median"]))
PY
tail -12 $O/trace.err | cut -c1-220
