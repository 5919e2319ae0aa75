#!/bin/bash
O=gpurun_out/${1:-final3}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
ensure_model('birdnet_v24')"
bash tools/gpu/tests.sh $1 | tail -3
for l in 3 1; do BN_COMPUTE_LANES=$l timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-ingest > $O/bench_cfg2_lanes$l.json 2> $O/bench_cfg2_lanes$l.err; done
timeout 900 python bench.py --steps 30 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
for f in cfg2_lanes3 cfg2_lanes1 cfg2; do python - <<PY
import json
d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
print("$f: value %.0f e2e %.0f pageable %s ms/step %.3f traffic %s" % (d["value"], d["e2e"]["value"], d.get("e2e_pageable") and round(d["e2e_pageable"]["value"]), d["ms_per_step"], d["roofline"].get("traffic")))
PY
done
