#!/bin/bash
O=gpurun_out/${1:-feab}; mkdir -p $O
bash tools/gpu/fe.sh $1
for m in 0 1; do
  BN_DISABLE_FE_FUSED=$m timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ingest > $O/bench_fused_off$m.json 2> $O/bench_off$m.err
  python - <<PY
import json
d=json.loads(open("$O/bench_fused_off$m.json").read().strip().splitlines()[-1])
print("BN_DISABLE_FE_FUSED=$m value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], "launches", d["gpu_launches"])
PY
done
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
