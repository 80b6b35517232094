#!/bin/bash
# two-phase focused GPU session (bounded: a hang of the fused kernel must not take the box)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_physics.py tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -q -x -k "twophase" > gpurun_out/pytest_tp.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tp.log; tail -25 gpurun_out/pytest_tp.log
for mode in 1 0; do
  CHIMP_TP_FUSED=$mode timeout 300 python scripts/measure_configs.py twophase > gpurun_out/tp_fused$mode.json 2> gpurun_out/tp_fused$mode.err; tail -1 gpurun_out/tp_fused$mode.json; tail -3 gpurun_out/tp_fused$mode.err
done
