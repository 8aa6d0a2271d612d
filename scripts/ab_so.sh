#!/bin/bash
# A/B two builds of the library on the GPU box: the in-tree one ("new") and tmp_ab/old.so ("old"), alternating.
set -u
mkdir -p gpurun_out
cp stringwars_b200/lib/libstringwars_b200.so /tmp/new.so
run() { # tag, extra bench args
  python bench.py --steps 30 --warmup 5 --cpu-seconds 1 $2 > gpurun_out/ab_so.json 2> gpurun_out/ab_so.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_so.json'))
print('%-4s %-28s value %.0f kernel %.0f GCUPS kernel_ms %.4f %s' % ('$1', '$2', d['value'], d['roofline']['kernel_gcups'], d['roofline']['kernel_ms'], d['config']['kernel']))
PY
}
for rep in 1 2; do
  for which in new old; do
    if [ $which = old ]; then cp tmp_ab/old.so stringwars_b200/lib/libstringwars_b200.so; else cp /tmp/new.so stringwars_b200/lib/libstringwars_b200.so; fi
    run $which ""
    run $which "--workload ascii256 --pairs 262144"
  done
done
cp /tmp/new.so stringwars_b200/lib/libstringwars_b200.so
