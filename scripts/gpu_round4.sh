#!/bin/bash
# one-GPU round session: all GPU tests + smoke, default bench (both arms), ncu of the two-phase kernels and launch lists
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-2500; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-900
# two-phase kernels: full ncu sections of one moment pass and one collide pass, and the launch list of a few steps
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"phaseMoments|twoPhaseCollide" -s 6 -c 2 -o gpurun_out/tp_kernels_384 -f python scripts/measure_configs.py twophase > gpurun_out/ncu_tp2.log 2>&1; tail -2 gpurun_out/ncu_tp2.log
TWOPHASE_SIZE=256 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/tp_launches_256.csv python scripts/measure_configs.py twophase > /dev/null 2>&1; tail -8 gpurun_out/tp_launches_256.csv | cut -c1-200
