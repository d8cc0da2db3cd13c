#!/bin/bash
# usage: scripts/build_variant.sh <tag> <size> [extra nvcc flags...]
# Builds element-crush-gym_b200/lib/libecg_<tag>.so: the kernels of ONE board size recompiled from the working tree
# with the extra flags, every other object reused from the regular build (run element-crush-gym_b200/build.py first).
set -e
tag=$1; size=$2; shift; shift
P=$(dirname "$0")/../element-crush-gym_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --diag-suppress 177 \
  -DECG_SIZE=$size "$@" -c $P/csrc/ecg_shape_kernels.cu -o $P/build/shape_${size}_$tag.o
objs=""
for n in 4 5 6 7 8 9 10 11 12 13 14 15 16; do
  if [ $n = $size ]; then objs="$objs $P/build/shape_${size}_$tag.o"; else objs="$objs $P/build/shape_$n.o"; fi
done
nvcc -shared -o $P/lib/libecg_$tag.so $P/build/api.o $objs -gencode arch=compute_100a,code=sm_100a
echo built libecg_$tag.so
