#!/bin/bash
# Build an experimental variant of the library with extra nvcc flags into variants/<name>/magpy_b200
# (git-ignored; travels to the GPU box).  Usage: scripts/build_variant.sh <name> [-DFLAG ...]
#   then on the box:  MB_ROOT=variants/<name> python scripts/probe2.py
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
dst=$root/variants/$name/magpy_b200
mkdir -p $dst
cp $root/magpy_b200/*.py $root/magpy_b200/core*.so $dst/
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" \
    -shared -o $dst/libmagpy_b200.so $root/magpy_b200/csrc/magpy_b200.cu -lcudart
echo built $dst
