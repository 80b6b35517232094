#!/bin/bash
# bench + ncu passes only
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; tail -1 gpurun_out/bench_256.json; tail -3 gpurun_out/bench_256.err
timeout 600 python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline --index table > gpurun_out/bench_256_table.json 2>> gpurun_out/bench_256.err; tail -1 gpurun_out/bench_256_table.json
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -1 gpurun_out/bench_512.json; tail -3 gpurun_out/bench_512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --size 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:collideStream -s 5 -c 2 -o gpurun_out/prof_collide -f python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
