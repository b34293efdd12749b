#!/bin/bash
# tools/build_variant.sh <tag> <nvcc defines...>: rebuilds the n_fft-2048 configuration (64 threads x 17 bins)
# with extra defines and links it with the other objects into _lib/variants/libsot_<tag>.so
# (use with SOT_B200_LIBRARY=... python bench.py).  The default library must be built first.
set -e
TAG=$1; shift
PKG=/root/repo/1d-spectral-optimal-transport_b200
mkdir -p $PKG/_lib/variants /root/repo/gpurun_out/scratch/sot_variant_$TAG
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
    -c $PKG/csrc/sot_cfg_64_17_1096_1.cu -o /root/repo/gpurun_out/scratch/sot_variant_$TAG/cfg.o 2> /root/repo/gpurun_out/scratch/sot_variant_$TAG/ptxas.log
OBJS=$(ls $PKG/build/*.o | grep -v sot_cfg_64_17_1096_1.o)
nvcc -shared -o $PKG/_lib/variants/libsot_$TAG.so /root/repo/gpurun_out/scratch/sot_variant_$TAG/cfg.o $OBJS -gencode arch=compute_100a,code=sm_100a
grep -A1 "ILi64ELi17ELi1096ELi1ELb1ELb0ELi2ELi[01]ELi0" /root/repo/gpurun_out/scratch/sot_variant_$TAG/ptxas.log | grep -E "Used|spill" | head -4
