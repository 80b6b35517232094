#!/bin/bash
set -x
mkdir -p gpurun_out
for v in tp4 tp5 tp4pm4 tp256; do
  CHIMP_LIB=$PWD/badchimp-cpp_b200/build/variants/libchimp_$v.so timeout 300 python scripts/measure_configs.py twophase 2> gpurun_out/tp_$v.err | tee gpurun_out/tp_$v.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('TP $v MLUPS %.0f frac %.3f ms %.3f'%(d['MLUPS'],d['frac_of_measured_hbm'],d['ms_per_step']))"
done
CHIMP_LIB=$PWD/badchimp-cpp_b200/build/variants/libchimp_tp4.so timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'phaseMoments|twoPhaseCollide|fluxForce' -s 9 -c 6 --csv --log-file gpurun_out/tp_launches.csv python scripts/measure_configs.py twophase > /dev/null 2>&1
cat gpurun_out/tp_launches.csv | tail -20 | cut -c1-250
