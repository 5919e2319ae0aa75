#!/bin/bash
# N-GPU bench lines (one rank per GPU, torchrun): cfg2 weak scaling and cfg5 strong scaling
O=gpurun_out/${1:-multiN}; N=${2:-2}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
ensure_model('birdnet_v24')"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --no-cpu-baseline > $O/bench_cfg2_${N}gpu.json 2> $O/bench_cfg2_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config 5 --no-cpu-baseline > $O/bench_cfg5_${N}gpu.json 2> $O/bench_cfg5_${N}gpu.err
for f in cfg2 cfg5; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${f}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$f N=$N: value %.0f e2e %.0f pageable %s" % (d["value"], d["e2e"]["value"], d.get("e2e_pageable") and round(d["e2e_pageable"]["value"])))
except Exception as e: print("$f failed", e)
PY
tail -2 $O/bench_${f}_${N}gpu.err; done
