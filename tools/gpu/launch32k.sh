#!/bin/bash
# ncu launch lists + live stage tables of the 32 kHz configurations (cfg3 v3.0 B=512, cfg4 Perch B=256)
O=gpurun_out/${1:-launch32k}; mkdir -p $O
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
for f in ('birdnet_v30','perch_v2'): ensure_model(f)"
for c in 3 4; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-ingest > $O/bench_short_cfg$c.json 2> $O/bench_short_cfg$c.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/ncu_launches_cfg$c.csv \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-ingest > $O/ncu_launches_cfg$c.log 2>&1
  ls -la $O/ncu_launches_cfg$c.csv
done
