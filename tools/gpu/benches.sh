#!/bin/bash
# bench lines of every BASELINE.json configuration at HEAD (1 GPU); logs under gpurun_out/$1
O=gpurun_out/${1:-benches}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
for f in ('birdnet_v24','birdnet_v30','perch_v2'): ensure_model(f)"
if [ -n "$LANES_AB" ]; then
BN_COMPUTE_LANES=3 timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-ingest > $O/bench_cfg2_lanes3.json 2> $O/bench_cfg2_lanes3.err
BN_COMPUTE_LANES=1 timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-ingest > $O/bench_cfg2_lanes1.json 2> $O/bench_cfg2_lanes1.err
fi
timeout 900 python bench.py --steps 30 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --config 1 --steps 60 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 900 python bench.py --config 3 --steps 12 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 900 python bench.py --config 4 --steps 12 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
timeout 900 python bench.py --config 5 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for f in cfg2 cfg1 cfg3 cfg4 cfg5 reference_arm; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
    print("$f: value %.0f e2e %.0f pageable %s ms/step %.3f" % (d["value"], d["e2e"]["value"], d.get("e2e_pageable") and round(d["e2e_pageable"]["value"]), d["ms_per_step"]))
except Exception as e: print("$f failed", e)
PY
done
