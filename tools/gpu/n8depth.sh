#!/bin/bash
O=gpurun_out/${1:-n8depth}; D=${2:-3}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
ensure_model('birdnet_v24')"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --no-cpu-baseline --no-ingest --pipeline-depth $D > $O/bench_cfg2_8gpu_depth$D.json 2> $O/bench_cfg2_8gpu_depth$D.err
python - <<PY
import json
d=json.loads(open("$O/bench_cfg2_8gpu_depth$D.json").read().strip().splitlines()[-1])
print("depth $D N=8: value %.0f e2e %.0f pageable %.0f" % (d["value"], d["e2e"]["value"], d["e2e_pageable"]["value"]), d["config"]["e2e_run_values"])
PY
