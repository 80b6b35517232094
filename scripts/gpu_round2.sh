#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python scripts/measure_configs.py d2q9 d3q27 trt twophase > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -1 gpurun_out/bench_512.json; tail -3 gpurun_out/bench_512.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench512.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:collideStream -s 6 -c 1 -o gpurun_out/prof_final_compact_512 -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
