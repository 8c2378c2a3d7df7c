#!/bin/bash
# ncu --set full of the cluster / small / implicit kernels (one short launch each)
set -u
mkdir -p gpurun_out
prof() {  # name regex N mode R steps
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -f -o gpurun_out/$1 \
        python scripts/probe_one.py $3 $4 $5 $6 > gpurun_out/ncu_$1.log 2>&1
    tail -1 gpurun_out/ncu_$1.log
}
prof heun_cluster64 heun_cluster 64 heun 9472 200
prof heun_cluster8 heun_cluster 8 heun 131072 400
prof heun_small2 heun_small 2 heun 524288 1000
prof imid_single imid_single 1 implicit 524288 200
prof imid_small2 imid_small 2 implicit 262144 200
ls -la gpurun_out/*.ncu-rep
