#!/bin/bash
# tools/build_variant.sh <tag> <config, e.g. 64_17_1096_1> <nvcc defines...>: rebuilds ONE kernel configuration with
# extra defines and links it with the other objects of the default build into _lib/variants/libsot_<tag>.so
# (use with SOT_B200_LIBRARY=... python bench.py; tools/bench_variants.sh runs them all).  Build the default first.
set -e
TAG=$1; CFG=$2; shift; shift
PKG=/root/repo/1d-spectral-optimal-transport_b200
TMP=/root/repo/gpurun_out/scratch/sot_variant_$TAG
mkdir -p $PKG/_lib/variants $TMP
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
    -c $PKG/csrc/sot_cfg_$CFG.cu -o $TMP/cfg.o 2> $TMP/ptxas.log
OBJS=$(ls $PKG/build/*.o | grep -v sot_cfg_$CFG.o)
nvcc -shared -o $PKG/_lib/variants/libsot_$TAG.so $TMP/cfg.o $OBJS -gencode arch=compute_100a,code=sm_100a
grep -A2 "ELb1ELb0ELi2ELi[01]ELi0" $TMP/ptxas.log | grep -E "Used|spill" | head -4
