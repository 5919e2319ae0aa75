#!/bin/bash
# round-2 GPU call A: tests at HEAD, staging microbench, bench lines of configs 2/1/3/4, C e2e driver
set -u
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
nproc >> $O/gpus.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" >> $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 tools/_build/pcie_bench > $O/pcie_bench_1gpu.json 2> $O/pcie_bench.err; cat $O/pcie_bench_1gpu.json
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -c 1500 $O/bench_cfg2.json
M=models/birdnet_v24_seed0.onnx
for pin in 0 1; do timeout 300 tools/_build/e2e_driver --model $M --mode ctx --depth 5 --batches 60 --pinned $pin >> $O/e2e_c_driver.jsonl 2>> $O/e2e_c_driver.err; done
timeout 300 tools/_build/e2e_driver --model $M --mode pool --depth 3 --batches 40 --pinned 1 --range 1 >> $O/e2e_c_driver.jsonl 2>> $O/e2e_c_driver.err
cat $O/e2e_c_driver.jsonl
timeout 600 python bench.py --config 1 --steps 40 --no-cpu-baseline > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 900 python bench.py --config 3 --steps 10 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 900 python bench.py --config 4 --steps 10 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err
for c in 1 3 4; do echo "cfg$c: $(head -c 400 $O/bench_cfg$c.json)"; tail -3 $O/bench_cfg$c.err; done
