#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; tail -1 gpurun_out/bench_256.json
timeout 600 python bench.py --size 256 --steps 50 --warmup 5 --no-cpu-baseline --index table > gpurun_out/bench_256_table.json 2>> gpurun_out/bench_256.err; tail -1 gpurun_out/bench_256_table.json
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -1 gpurun_out/bench_512.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --size 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:collideStream -s 5 -c 2 -o gpurun_out/prof_collide -f python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
