#!/bin/bash
# usage: gpu_r2_bench_n.sh N [extra bench args]  -- strong-scaling bench line on N GPUs + phase trace
N=${1:-2}; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 10 "$@" > gpurun_out/bench_strong_${N}.json 2> gpurun_out/bench_strong_${N}.err; echo "bench exit $?"; tail -c 4500 gpurun_out/bench_strong_${N}.json; grep -v "^W\|^\[W\|warn" gpurun_out/bench_strong_${N}.err | tail -5
CHIMP_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 10 --no-parity --no-weak "$@" > gpurun_out/bench_trace_${N}.json 2> gpurun_out/bench_trace_${N}.err; grep "chimp trace" gpurun_out/bench_trace_${N}.err | grep " 100 steps" | cut -c1-330
