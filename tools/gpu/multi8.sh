#!/bin/bash
# 8-GPU box: staging roofline, C-driver e2e (no Python), pool (cfg5 shape), torchrun bench cfg5 and cfg2
O=gpurun_out/${1:-multi8}; mkdir -p $O
N=${2:-8}
python -c "
import sys; sys.path.insert(0,'rust-birdnet-onnx_b200')
from birdnet_b200.modelgen.make_models import ensure_model
ensure_model('birdnet_v24')"
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)" >> $O/gpus.txt; free -g | head -2 >> $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 tools/_build/pcie_bench > $O/pcie_bench_8gpu.json 2> $O/pcie_bench.err; cat $O/pcie_bench_8gpu.json
M=models/birdnet_v24_seed0.onnx
D8=0,1,2,3,4,5,6,7
timeout 300 tools/_build/e2e_driver --model $M --mode ctx --devices $D8 --depth 4 --batches 48 --pinned 1 --reps 3 >> $O/e2e_c_driver.jsonl 2>> $O/e2e_c_driver.err
timeout 300 tools/_build/e2e_driver --model $M --mode ctx --devices $D8 --depth 4 --batches 48 --pinned 0 --reps 3 >> $O/e2e_c_driver.jsonl 2>> $O/e2e_c_driver.err
timeout 300 tools/_build/e2e_driver --model $M --mode ctx --devices 0,1,2,3 --depth 4 --batches 48 --pinned 1 --reps 3 >> $O/e2e_c_driver.jsonl 2>> $O/e2e_c_driver.err
timeout 300 tools/_build/e2e_driver --model $M --mode pool --devices $D8 --depth 3 --batches 14 --pinned 1 --range 1 --reps 3 >> $O/e2e_c_driver.jsonl 2>> $O/e2e_c_driver.err
timeout 300 tools/_build/e2e_driver --model $M --mode pool --devices $D8 --depth 3 --batches 14 --pinned 0 --range 1 --reps 3 >> $O/e2e_c_driver.jsonl 2>> $O/e2e_c_driver.err
cat $O/e2e_c_driver.jsonl; tail -3 $O/e2e_c_driver.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config 5 --no-cpu-baseline > $O/bench_cfg5_8gpu.json 2> $O/bench_cfg5_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --no-cpu-baseline > $O/bench_cfg2_8gpu.json 2> $O/bench_cfg2_8gpu.err
for f in cfg5 cfg2; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${f}_8gpu.json").read().strip().splitlines()[-1])
    print("$f N=8: value %.0f e2e %.0f pageable %.0f ingest %s" % (d["value"], d["e2e"]["value"], d["e2e_pageable"]["value"], d["ingest_pcm16"] and round(d["ingest_pcm16"]["value"])))
except Exception as e: print("$f failed", e)
PY
tail -2 $O/bench_${f}_8gpu.err; done
# the in-process pool on 2 real devices + the gloo-free multi-GPU tests
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pool or cfg5" > $O/pytest_pool.log 2>&1; tail -3 $O/pytest_pool.log
