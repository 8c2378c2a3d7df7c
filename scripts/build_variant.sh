#!/bin/bash
# Build an experimental variant of the library into variants/<name>/magpy_b200 (git- and gpurun-ignored... the
# latter only if you remove `variants/` from .gpurunignore for the experiment).
#   scripts/build_variant.sh <name> <unit.cu>[,<unit2.cu>...] [extra nvcc flags for those units ...]
#   on the box:  MB_ROOT=variants/<name> python scripts/probe2.py f32p
set -e
name=$1; unit=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/magpy_b200/csrc
dst=$root/variants/$name/magpy_b200
mkdir -p $dst/obj
cp $root/magpy_b200/*.py $root/magpy_b200/core*.so $dst/
objs=""
for u in magpy_b200 comm heun_single heun_single_balanced heun_single_split imid_single small_heun small_imid cluster cluster_mma cluster_mma_imid cluster_big service dom; do
    if [[ ",$unit," == *",$u.cu,"* ]]; then
        nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c -o $dst/obj/$u.o $src/$u.cu
        objs="$objs $dst/obj/$u.o"
    else
        objs="$objs $src/build/$u.o"
    fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $dst/libmagpy_b200.so $objs -lcudart -ldl
echo built $dst
