#!/bin/bash
# round 2, multi-GPU pass after the K1 rework (gpurun --gpus 8): the driver's launch line at N = 1, 2, 4, 8 (strong scaling, the
# library's own NCCL communicator), config 4 at N = 8, the two-rank NCCL test.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
echo "== two-rank NCCL test"
timeout 600 python -m pytest tests/test_sharding_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_sharding_multi.log
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 2>&1 | tail -1 | tee gpurun_out/r02_scale_n1.json
for n in 2 4 8; do
  if [ "$n" -le "$NG" ]; then
    echo "== bench N=$n"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
        bench.py --gpus $n 2>&1 | grep '^{' | tail -1 | tee gpurun_out/r02_scale_n$n.json
  fi
done
if [ "$NG" -ge 8 ]; then
  echo "== bench c4 N=8"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 8 --workload c4 2>&1 | grep '^{' | tail -1 | tee gpurun_out/r02_scale_c4_n8.json
fi
