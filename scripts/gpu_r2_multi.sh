#!/bin/bash
# usage: gpu_r2_multi.sh N  -- N-GPU parity check (peer halos) + strong-scaling bench + phase trace
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/multi_check_${N}.log 2>&1; echo "multi check exit $?"; grep -v "^W\|^\[W\|warn" gpurun_out/multi_check_${N}.log | tail -16 | cut -c1-420
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_strong_${N}.json 2> gpurun_out/bench_strong_${N}.err; echo "bench exit $?"; tail -c 4500 gpurun_out/bench_strong_${N}.json; grep -v "^W\|^\[W\|warn" gpurun_out/bench_strong_${N}.err | tail -5
CHIMP_TRACE=1 timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 10 --no-parity --no-weak > gpurun_out/bench_trace_${N}.json 2> gpurun_out/bench_trace_${N}.err; grep "chimp trace" gpurun_out/bench_trace_${N}.err | tail -$((3*N)) | cut -c1-330
