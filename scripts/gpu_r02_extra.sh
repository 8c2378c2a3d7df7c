#!/bin/bash
# round 2, extra evidence after the K1 rework: ncu full captures of K1s (config 1 shape) and of the implicit dimer kernel
# (config 2 shape), BASELINE config 3 at full physical length and the config 5 convergence sweep with the new stream layout.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_r02_extra.sh'
set -u
mkdir -p gpurun_out
echo "== ncu full: heun_single_split (1000 members x 4000 steps)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heun_single_split -c 1 \
    -f -o gpurun_out/r02_heun_single_split python scripts/probe_one.py 1 heun 1000 4000 > gpurun_out/r02_ncu_k1s.log 2>&1
tail -2 gpurun_out/r02_ncu_k1s.log
echo "== ncu full: imid_small (10,000 dimers x 100 steps)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:imid_small -c 1 \
    -f -o gpurun_out/r02_imid_small2 python scripts/probe_one.py 2 implicit 10000 100 > gpurun_out/r02_ncu_imid_small2.log 2>&1
tail -2 gpurun_out/r02_ncu_imid_small2.log
echo "== config 3 at full length (12 nm, then 6 nm)"
timeout 600 python scripts/run_config3_full.py 2>&1 | tail -1 | tee gpurun_out/r02_config3_full.jsonl
timeout 600 python scripts/run_config3_full.py 6e-9 2>&1 | tail -1 | tee -a gpurun_out/r02_config3_full.jsonl
echo "== config 5 convergence sweep"
timeout 900 python scripts/run_config5_convergence.py 2>&1 | grep '^{' | tee gpurun_out/r02_config5_convergence.jsonl | cut -c1-300
