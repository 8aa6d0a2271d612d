#!/bin/bash
# Round-2 evidence sweep, one gpurun call on one GPU: launch list of the default bench command, one `ncu --set full` capture
# per kernel the bench line's gates name (summarised on the box: reports exceed gpurun_out's 64 MiB cap), sanitizer run.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2_smi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r2_smi.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --no-gates > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_utf8.csv python scripts/utf8_probe.py 100000 > /dev/null 2>&1
python - <<PY > gpurun_out/r2_launches_utf8.txt
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_utf8.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:70]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[vi].replace(",", ""))
for k, (c, t) in agg.items(): print(f"{c:4d} launches, {t / c / 1000:9.1f} us each  {k}")
PY
cat gpurun_out/r2_launches_utf8.txt
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    $NCU -k regex:$regex -s $skip -c 1 -o gpurun_out/$name "$@" > /dev/null 2>&1
    python scripts/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.txt > /dev/null 2>&1
    rm -f gpurun_out/$name.ncu-rep
}
cap r2_bitpar_k2_ascii64 lev_bitpar_kernel 3 python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --no-gates
cap r2_bitpar_k8_ascii256 lev_bitpar_kernel 3 python bench.py --workload ascii256 --steps 2 --warmup 3 --cpu-seconds 1 --no-gates
cap r2_load_phase_ascii64 load_phase 2 python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --no-gates
cap r2_load_phase_ascii256 load_phase 2 python bench.py --workload ascii256 --steps 2 --warmup 3 --cpu-seconds 1 --no-gates
cap r2_short_w4_words lev_short_kernel 3 python scripts/words_probe.py
cap r2_cross_k1_words lev_bitpar_cross_kernel 3 python scripts/words_probe.py
cap r2_utf8_rename utf8_rename_kernel 2 python scripts/utf8_probe.py 100000
cap r2_wavefront_long align_wavefront_kernel 0 python -m pytest tests/test_gpu_boundary.py -q -m gpu -k longer_than_shared
python scripts/words_probe.py > gpurun_out/r2_words_probe.log 2>&1
python scripts/utf8_probe.py > gpurun_out/r2_utf8_probe.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_memcheck.log
tail -3 gpurun_out/r2_memcheck.log
