#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: K1s (heun_single_split.cu: named barriers, shared-memory ring),
# the capacity-free implicit kernel (cluster_big.cu: two global iterate buffers, CTA-uniform loop with per-member state),
# and the reworked K1 / K1b.   gpurun --timeout 2400 -- 'bash scripts/gpu_r02_sanitize2.sh'
set -u
mkdir -p gpurun_out
K='capacity_free and not 130 or variants_are_bit or balanced_persistent'
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "$K" 2>&1 | tail -8 | tee gpurun_out/r02_compute_sanitizer_memcheck_late.log
echo "== racecheck"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "capacity_free and (20-True or 33-True) or variants_are_bit" 2>&1 | tail -8 | tee gpurun_out/r02_compute_sanitizer_racecheck_late.log
echo "== synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "capacity_free and (20-True or 33-True) or variants_are_bit" 2>&1 | tail -8 | tee gpurun_out/r02_compute_sanitizer_synccheck_late.log
