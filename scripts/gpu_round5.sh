#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/measure_transfers.py > gpurun_out/transfers.json 2> gpurun_out/transfers.err; cat gpurun_out/transfers.json; tail -2 gpurun_out/transfers.err
CHIMP_LIB=$PWD/badchimp-cpp_b200/build/variants/libchimp_tpmb5.so timeout 300 python scripts/measure_configs.py twophase 2>&1 | tail -1 | cut -c1-330
timeout 300 python scripts/measure_configs.py twophase 2>&1 | tail -1 | cut -c1-330
TWOPHASE_SIZE=256 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"phaseMoments|twoPhaseCollide|fluxForce" -c 30 --csv --log-file gpurun_out/tp_launches_256.csv python scripts/measure_configs.py twophase > /dev/null 2>&1; tail -6 gpurun_out/tp_launches_256.csv | cut -c1-260
