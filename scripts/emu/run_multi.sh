#!/bin/bash
# DEVELOPMENT TOOL (see include/cuda_runtime.h): N ranks as N processes on the host model.  Every rank's cudaMalloc
# comes out of one shared-memory arena mapped at the same address (CHIMP_EMU_ARENA), so peer pointers and "IPC handles"
# mean the same memory in all ranks; torch.distributed runs over gloo.  Run after scripts/emu/run.sh has built a copy:
#   scripts/emu/run_multi.sh N /root/repo/scripts/emu/bench_dry_run.py --gpus N [bench.py arguments]
#   scripts/emu/run_multi.sh N tests/multi_gpu_check.py
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=${EMU_ROOT:-/tmp/chimp_emu/repo}
N=$1; shift
ARENA=/dev/shm/chimp_emu_arena_$$
rm -f "$ARENA"
trap 'rm -f "$ARENA"' EXIT
cd "$ROOT"
PORT=$((20000 + RANDOM % 20000))
CHIMP_EMU=1 CHIMP_EMU_ARENA="$ARENA" PYTHONPATH="$HERE:$PYTHONPATH" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port "$PORT" "$@"
