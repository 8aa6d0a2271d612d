#!/bin/bash
# A/B several builds of the library on the GPU box: the in-tree one ("base") and every variants/*.so, two rounds each.
# usage: ab_variants.sh [bench args for a second workload]
set -u
mkdir -p gpurun_out
cp stringwars_b200/lib/libstringwars_b200.so /tmp/base.so
run() { # tag, extra bench args
  python bench.py --steps 30 --warmup 5 --cpu-seconds 1 $2 > gpurun_out/ab_var.json 2> gpurun_out/ab_var.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_var.json'))
    print('%-12s %-38s value %.0f kernel %.0f GCUPS kernel_ms %.4f %s checksum %s' % ('$1', '$2', d['value'], d['roofline']['kernel_gcups'], d['roofline']['kernel_ms'], d['config']['kernel'], d.get('checksum')))
except Exception as e:
    print('$1', '$2', 'FAILED', e); print(open('gpurun_out/ab_var.err').read()[-600:])
PY
}
for rep in 1 2; do
  for so in /tmp/base.so variants/*.so; do
    tag=$(basename $so .so)
    cp $so stringwars_b200/lib/libstringwars_b200.so
    run $tag ""
    run $tag "--workload ascii256 --pairs 262144"
  done
done
cp /tmp/base.so stringwars_b200/lib/libstringwars_b200.so
