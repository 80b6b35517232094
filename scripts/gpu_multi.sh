#!/bin/bash
# N-GPU session: parity check against the undecomposed run, then the weak-scaling bench through torchrun exactly as the driver launches it
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; tail -6 gpurun_out/multi_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/multi_${N}_512.json 2> gpurun_out/multi_${N}_512.err; tail -1 gpurun_out/multi_${N}_512.json; tail -3 gpurun_out/multi_${N}_512.err
if [ "$2" == "with1" ]; then timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/multi_1_512.json 2> gpurun_out/multi_1_512.err; tail -1 gpurun_out/multi_1_512.json; fi
