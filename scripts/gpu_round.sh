#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/measure_transfers.py > gpurun_out/transfers.json 2> gpurun_out/transfers.err; python -c "
import json; d=json.load(open('gpurun_out/transfers.json')); print({k:(v if not isinstance(v,dict) else {a:round(b,4) for a,b in v.items()}) for k,v in d.items()})"; tail -2 gpurun_out/transfers.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e'], d['clocks'])"; tail -2 gpurun_out/bench_default.err
timeout 200 python scripts/measure_configs.py twophase 2>&1 | tail -1 | cut -c1-330
