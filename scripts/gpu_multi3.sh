#!/bin/bash
# N-GPU session: parity checks with peer memory, configs[3] with both transports
set -x
N=${1:-2}
mkdir -p gpurun_out
CHIMP_HALO=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_check_${N}_peer.log 2>&1; tail -4 gpurun_out/multi_check_${N}_peer.log | cut -c1-300
for mode in peer nccl; do
CHIMP_HALO=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 scripts/measure_multi.py twophase > gpurun_out/measure_tp_${N}_$mode.jsonl 2> gpurun_out/measure_tp_${N}_$mode.err; grep config gpurun_out/measure_tp_${N}_$mode.jsonl | cut -c1-500; tail -3 gpurun_out/measure_tp_${N}_$mode.err | cut -c1-300
done
