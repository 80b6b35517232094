#!/bin/bash
set -x
N=${1:-8}
mkdir -p gpurun_out
for mode in ${MODES:-peer nccl}; do
CHIMP_HALO=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/measure_multi.py twophase > gpurun_out/measure_tp_${N}_$mode.jsonl 2> gpurun_out/measure_tp_${N}_$mode.err; grep config gpurun_out/measure_tp_${N}_$mode.jsonl | cut -c1-700; tail -3 gpurun_out/measure_tp_${N}_$mode.err | cut -c1-300
done
