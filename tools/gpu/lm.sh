#!/bin/bash
O=gpurun_out/${1:-lm}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity_32k.py -m gpu -q -x > $O/pytest32k.log 2>&1; echo "pytest exit $?" >> $O/pytest32k.log; tail -4 $O/pytest32k.log
for c in 3 4; do timeout 600 python bench.py --config $c --steps 12 --no-cpu-baseline --no-ingest > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err
python - <<PY
import json
d=json.loads(open("$O/bench_cfg$c.json").read().strip().splitlines()[-1])
lm=[s for s in d["stages"] if s["stage"]=="logmel"]
print("cfg$c value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), "logmel", lm and (lm[0]["ms"], lm[0]["frac"]))
PY
done
