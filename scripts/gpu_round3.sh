#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python scripts/measure_configs.py ${CONFIGS:-twophase} > gpurun_out/configs2.jsonl 2> gpurun_out/configs2.err; cat gpurun_out/configs2.jsonl; tail -3 gpurun_out/configs2.err
