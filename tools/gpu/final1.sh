#!/bin/bash
# round-end evidence on one GPU: tests, profile set, every bench configuration, staging microbench, C driver
O=${1:-final1}
bash tools/gpu/tests.sh $O
bash tools/gpu/profile.sh $O
bash tools/gpu/benches.sh $O
timeout 300 tools/_build/pcie_bench > gpurun_out/$O/pcie_bench_1gpu.json 2> gpurun_out/$O/pcie_bench.err; tail -3 gpurun_out/$O/pcie_bench_1gpu.json | cut -c1-300
M=models/birdnet_v24_seed0.onnx
for pin in 1 0; do timeout 300 tools/_build/e2e_driver --model $M --mode ctx --depth 5 --batches 60 --pinned $pin --reps 3 >> gpurun_out/$O/e2e_c_driver_1gpu.jsonl 2>> gpurun_out/$O/e2e_c_driver.err; done
timeout 300 tools/_build/e2e_driver --model $M --mode pool --depth 5 --batches 60 --pinned 1 --range 1 --reps 3 >> gpurun_out/$O/e2e_c_driver_1gpu.jsonl 2>> gpurun_out/$O/e2e_c_driver.err
cat gpurun_out/$O/e2e_c_driver_1gpu.jsonl | cut -c1-300
