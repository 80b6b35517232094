#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-256_3 256_4 128_6 128_8}; do
  for idx in ${FORMS:-compact table}; do
    CHIMP_LIB=$PWD/badchimp-cpp_b200/build/variants/libchimp_$v.so timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --index $idx > gpurun_out/sweep_${v}_$idx.json 2> gpurun_out/sweep_${v}_$idx.err
    python - <<PY || tail -2 gpurun_out/sweep_${v}_$idx.err
import json
d=json.loads(open('gpurun_out/sweep_${v}_$idx.json').read().strip().splitlines()[-1])
print('SWEEP ${v} $idx MLUPS %.0f frac %.3f clocks %s idxB %.1f irr %.4f e2e %.0f'%(d['value'],d['roofline']['frac'],d['clocks']['sm_mhz'],d['config']['index_bytes_per_node'],d['config']['irregular_tile_fraction'],d['e2e']['value'] or 0))
PY
  done
done
