#!/bin/bash
# usage: scripts/build_variant.sh <tag> <size> [extra nvcc flags...]
# Builds element-crush-gym_b200/lib/libecg_<tag>.so holding ONLY board size <size>, its kernels compiled from the
# working tree with the extra flags (A/B experiments: ECG_LIB=.../libecg_<tag>.so selects it).
set -e
tag=$1; size=$2; shift; shift
P=$(dirname "$0")/../element-crush-gym_b200
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --diag-suppress 177"
nvcc $F -DECG_SIZE=$size "$@" -c $P/csrc/ecg_shape_kernels.cu -o $P/build/shape_${size}_$tag.o
[ -f $P/build/api_only_$size.o ] && [ $P/build/api_only_$size.o -nt $P/csrc/ecg_api.cu ] || \
  nvcc $F -DECG_ONLY_SIZE=$size -c $P/csrc/ecg_api.cu -o $P/build/api_only_$size.o
nvcc -shared -o $P/lib/libecg_$tag.so $P/build/api_only_$size.o $P/build/shape_${size}_$tag.o -gencode arch=compute_100a,code=sm_100a
echo built libecg_$tag.so
