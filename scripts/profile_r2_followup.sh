#!/bin/bash
# Round-2 follow-up sweep (one GPU): everything behind the profiles/ files added after profile_r2.sh — the code-point engine's
# rename pass and cross-product path, the ragged length classes, sanitizers over the kernels changed last.
#   bash scripts/profile_r2_followup.sh          (outputs land in gpurun_out/)
set -u
mkdir -p gpurun_out
summary() { # launch list csv -> per-kernel count / mean / min / max
python - "$1" <<PY
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")) / 1000)
for k, v in agg.items(): print(f"{len(v):4d} launches, mean {sum(v)/len(v):9.1f} us  min {min(v):8.1f} max {max(v):8.1f}  {k}")
PY
}
cap() { # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/$name "$@" > /dev/null 2>&1
    python scripts/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.txt > /dev/null 2>&1
    rm -f gpurun_out/$name.ncu-rep
}
# config 3 (100 k UTF-8 line pairs): per-call times, launch list, captures of the rename pass and of the two big ragged classes
python scripts/utf8_probe.py > gpurun_out/r2_utf8_probe.log 2>&1; cat gpurun_out/r2_utf8_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_utf8_after.csv python scripts/utf8_probe.py 100000 > /dev/null 2>&1
summary gpurun_out/r2_launches_utf8_after.csv | tee gpurun_out/r2_launches_utf8_after.txt
cap r2_utf8_rename utf8_rename_kernel 2 python scripts/utf8_probe.py 100000
cap r2_bitpar_k8_ragged_lines lev_bitpar_kernel 3 python scripts/utf8_probe.py 100000
cap r2_coop_g8k2_ragged_lines lev_bitpar_coop_kernel 1 python scripts/utf8_probe.py 100000
# the harness's shape for the code-point engine
python scripts/utf8_cross_probe.py 2>&1 | tee gpurun_out/r2_utf8_cross_probe.log
# sanitizers over the kernels changed last (rename tiers, batch dictionary, forked class streams)
SWB_PROBE_ONLY=utf8 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2_memcheck_changed.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_memcheck_changed.log
SWB_PROBE_ONLY=utf8 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2_racecheck_changed.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_racecheck_changed.log
tail -n 3 gpurun_out/r2_memcheck_changed.log; tail -n 3 gpurun_out/r2_racecheck_changed.log
