#!/bin/bash
# round 2, final single-GPU pass after the K1 rework (37-instruction step, new packed field layout, K1s): smoke, bench (both
# arms, all workloads), ncu launch list of the bench command, ncu --set full of the bench kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r02_final2.sh [--tests]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r02_gpu.csv 2>&1
if [ "${1:-}" == "--tests" ]; then
  echo "== pytest -m gpu"
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_final.log
fi
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_smoke.log
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_reference_n1.json
echo "== bench"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/r02_bench_n1.json
for wl in c4 c1 c2; do
  echo "== bench $wl"
  timeout 900 python bench.py --workload $wl 2>&1 | tail -1 | tee gpurun_out/r02_bench_${wl}_n1.json
done
echo "== ncu launch list (same command as the bench, shorter)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r02_bench_under_ncu.log | cut -c1-200
echo "== ncu full: heun_single (balanced persistent kernel, the bench kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heun_single_balanced -c 1 \
    -f -o gpurun_out/r02_heun_single_balanced python scripts/probe_one.py 1 heun 1000000 4000 > gpurun_out/r02_ncu_k1b.log 2>&1
tail -2 gpurun_out/r02_ncu_k1b.log
echo "== ncu full: heun_single_split (K1s, config 1 shape: 1000 members x 4000 steps)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heun_single_split -c 1 \
    -f -o gpurun_out/r02_heun_single_split python scripts/probe_one.py 1 heun 1000 4000 > gpurun_out/r02_ncu_k1s.log 2>&1
tail -2 gpurun_out/r02_ncu_k1s.log
ls -la gpurun_out | tail -12
