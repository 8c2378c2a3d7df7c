#!/bin/bash
# second GPU pass of the round: everything after the DMMA cluster kernel went in
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== peaks"
python -c "
import sys; sys.path.insert(0,'.')
import magpy_b200.core as c
print('dfma peak', c.fp64_peak(0)); print('dmma peak', c.fp64_mma_peak(0))" | tee gpurun_out/peaks.log
echo "== probe"
python scripts/probe_mma.py 2>&1 | tee gpurun_out/probe_mma.log
echo "== configs"
python scripts/run_configs.py 2>&1 | tee gpurun_out/configs_n1.jsonl
echo "== sanitizer (racecheck, memcheck) on the DMMA cluster tests"
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_parity_gpu.py -x -q -k "cluster_mma and (8-False or 24 or 64-True)" 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck_mma.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -x -q -k "cluster_mma and (8-False or 24 or 64-True)" 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck_mma.log
