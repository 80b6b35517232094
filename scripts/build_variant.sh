#!/bin/bash
# usage: build_variant.sh <tag> [extra nvcc -D flags...]  -> badchimp-cpp_b200/build/variants/libchimp_<tag>.so
TAG=$1; shift
OUT=badchimp-cpp_b200/build/variants/libchimp_${TAG}.so
mkdir -p badchimp-cpp_b200/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
  "$@" -o $OUT badchimp-cpp_b200/csrc/engine.cu badchimp-cpp_b200/csrc/kernels.cu 2>&1 | grep -E "error" ; echo built $OUT
