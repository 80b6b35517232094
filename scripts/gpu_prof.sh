#!/bin/bash
set -x
mkdir -p gpurun_out
for cfg in "128_8 compact" "256_3 table"; do set -- $cfg
CHIMP_LIB=$PWD/badchimp-cpp_b200/build/variants/libchimp_$1.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:collideStream -s 8 -c 1 -o gpurun_out/prof_$1_$2 -f python bench.py --size 384 --steps 6 --warmup 3 --no-cpu-baseline --index $2 > gpurun_out/ncu_$1_$2.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
