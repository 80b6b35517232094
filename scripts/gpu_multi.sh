#!/bin/bash
# N-GPU session: parity checks (incl. two-phase) + configs[3] / configs[4] measurements
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; tail -8 gpurun_out/multi_check_$N.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/measure_multi.py $2 > gpurun_out/measure_multi_$N.jsonl 2> gpurun_out/measure_multi_$N.err; grep config gpurun_out/measure_multi_$N.jsonl | cut -c1-600; tail -5 gpurun_out/measure_multi_$N.err | cut -c1-300
