#!/bin/bash
# tests + bench cfg2 + fused-vs-layered check; logs under gpurun_out/$1
O=gpurun_out/${1:-full}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 600 python tools/mbconv_check.py --batch 8 --times-batch 256 > $O/mbconv_check.log 2>&1; tail -12 $O/mbconv_check.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; python - <<PY
import json
d=json.load(open("$O/bench_cfg2.json"))
print("cfg2 value %.0f e2e %.0f pageable %.0f ms/step %.3f launches %d" % (d["value"], d["e2e"]["value"], d["e2e_pageable"]["value"], d["ms_per_step"], d["gpu_launches"]/d["steps"]))
for k in d["kernel_classes"][:8]: print("  ", k)
PY
