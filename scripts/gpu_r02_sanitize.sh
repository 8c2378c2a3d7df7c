#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (memcheck everywhere; racecheck on the shared-memory / named-barrier
# kernels).  gpurun --timeout 2400 -- 'bash scripts/gpu_r02_sanitize.sh'
set -u
mkdir -p gpurun_out
K='implicit_cluster_injected and (32-True or 64-True or 16-False or 12-True) or beyond_128 or per_member_material or per_member_parameters_do_not or balanced_persistent or variants_are_bit or adjugate'
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "$K" 2>&1 | tail -12 | tee gpurun_out/r02_compute_sanitizer_memcheck.log
echo "== racecheck (implicit DMMA cluster kernel: named barriers per column tile, one moment buffer)"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "implicit_cluster_injected and (8-True or 12-True or 32-True)" 2>&1 | tail -12 | tee gpurun_out/r02_compute_sanitizer_racecheck.log
echo "== synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "implicit_cluster_injected and (12-True or 32-True) or balanced_persistent" 2>&1 | tail -12 | tee gpurun_out/r02_compute_sanitizer_synccheck.log
