#!/bin/bash
# round 2, pass A (2 GPUs): sharding / communicator tests, bench at N=1 and N=2 through the driver's launch line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_gpu.csv 2>&1
echo "== sharding tests"
timeout 600 python -m pytest tests/test_sharding_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_sharding.log
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 2>&1 | tail -2 | tee gpurun_out/r02_bench_n1.json
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 2>&1 | tail -4 | tee gpurun_out/r02_bench_n2.json
