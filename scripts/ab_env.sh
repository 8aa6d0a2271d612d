#!/bin/bash
# A/B on the GPU box: default bench line under each "VAR=value[,VAR=value]" setting given ("-" = defaults).
# usage: ab_env.sh [--pytest] setting...
set -u
mkdir -p gpurun_out
if [ "${1:-}" = "--pytest" ]; then shift
  ( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -2 gpurun_out/pytest_gpu.log
fi
for setting in "$@"; do
  envs=$(echo "$setting" | tr ',' ' '); [ "$setting" = "-" ] && envs=""
  tag=$(echo "$setting" | tr -c 'A-Za-z0-9=\n' '_')
  env $envs python bench.py --steps 30 --warmup 5 --cpu-seconds 1 ${BENCH_ARGS:-} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$tag.json'))
    print('%-40s value %.0f kernel %.0f GCUPS kernel_ms %.4f  %s' % ('$setting', d['value'], d['roofline']['kernel_gcups'], d['roofline']['kernel_ms'], d['config']['kernel']))
except Exception as e:
    print('$setting', 'FAILED', e); print(open('gpurun_out/ab_$tag.err').read()[-800:])
PY
done
