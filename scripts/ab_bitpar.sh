#!/bin/bash
# A/B on the GPU box: the GPU suite, then the default bench line for each SWB_BITPAR_GRID value given.
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
for mode in "$@"; do
  SWB_BITPAR_GRID=$mode python bench.py --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/ab_grid$mode.json 2> gpurun_out/ab_grid$mode.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_grid$mode.json'))
print('grid mode $mode: value %.0f kernel %.0f GCUPS kernel_ms %.4f' % (d['value'], d['roofline']['kernel_gcups'], d['roofline']['kernel_ms']))
PY
done
