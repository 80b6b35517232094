#!/bin/bash
# usage: build_variant.sh <block> <minblocks> [extra nvcc flags...]  -> badchimp-cpp_b200/build/variants/libchimp_<block>_<minblocks>.so
B=$1; M=$2; shift 2
OUT=badchimp-cpp_b200/build/variants/libchimp_${B}_${M}.so
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
  -DCHIMP_BLOCK=$B -DCHIMP_MIN_BLOCKS=$M "$@" -o $OUT badchimp-cpp_b200/csrc/engine.cu badchimp-cpp_b200/csrc/kernels.cu 2>&1 | grep -E "error" ; echo built $OUT
