#!/bin/bash
# N-GPU session comparing the two halo transports: parity check + weak-scaling bench with peer stores and with NCCL
set -x
N=${1:-2}
mkdir -p gpurun_out
for mode in peer nccl; do
  CHIMP_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_check_${N}_$mode.log 2>&1; tail -4 gpurun_out/multi_check_${N}_$mode.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 100 --warmup 5 --halo $mode > gpurun_out/multi_${N}_512_$mode.json 2> gpurun_out/multi_${N}_512_$mode.err; tail -1 gpurun_out/multi_${N}_512_$mode.json; tail -3 gpurun_out/multi_${N}_512_$mode.err
done
if [ "$2" == "with1" ]; then timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/multi_1_512.json 2> gpurun_out/multi_1_512.err; tail -1 gpurun_out/multi_1_512.json; fi
