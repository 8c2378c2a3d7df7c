#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list of the bench command,
# ncu --set full of the dominant kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_reference.json
echo "== bench"
timeout 900 python bench.py 2>&1 | tail -3 | tee gpurun_out/bench.json
echo "== probe"
timeout 600 python scripts/probe.py 2>&1 | tail -20 | tee gpurun_out/probe.log
echo "== ncu launch list (same command as the bench, shorter)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log
echo "== ncu full: heun_single"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heun_single -c 1 \
    -f -o gpurun_out/heun_single python scripts/probe_one.py 1 heun 1000000 4000 > gpurun_out/ncu_heun_single.log 2>&1
tail -2 gpurun_out/ncu_heun_single.log
ls -la gpurun_out
