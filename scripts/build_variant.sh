#!/bin/bash
# usage: scripts/build_variant.sh <tag> [extra nvcc flags...]
# Builds element-crush-gym_b200/lib/libecg_<tag>.so: the 9x9 kernels recompiled from the working tree with the
# extra flags, every other object reused from the regular build (run element-crush-gym_b200/build.py first).
set -e
tag=$1; shift
P=$(dirname "$0")/../element-crush-gym_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --diag-suppress 177 \
  -DECG_SIZE=9 "$@" -c $P/csrc/ecg_shape_kernels.cu -o $P/build/shape_9_$tag.o
nvcc -shared -o $P/lib/libecg_$tag.so $P/build/api.o $P/build/shape_5.o $P/build/shape_6.o $P/build/shape_7.o \
  $P/build/shape_9_$tag.o $P/build/shape_12.o $P/build/shape_16.o -gencode arch=compute_100a,code=sm_100a
echo built libecg_$tag.so
