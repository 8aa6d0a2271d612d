set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_boundary.py tests/test_gpu_parity.py -x -q -m gpu -k "codepoint or utf8 or monotonic or long" 2>&1 | tail -5
python scripts/utf8_probe.py > gpurun_out/r2d_utf8_probe.log 2>&1; cat gpurun_out/r2d_utf8_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2d_launches_utf8.csv python scripts/utf8_probe.py 100000 > /dev/null 2>&1
python - <<PY > gpurun_out/r2d_launches_utf8.txt
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2d_launches_utf8.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:70]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[vi].replace(",", ""))
for k, (c, t) in agg.items(): print(f"{c:4d} launches, {t / c / 1000:9.1f} us each  {k}")
PY
cat gpurun_out/r2d_launches_utf8.txt
ncu --set full --clock-control none --import-source on -k regex:utf8_rename_kernel -s 2 -c 1 -o gpurun_out/r2d_utf8_rename python scripts/utf8_probe.py 100000 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2d_utf8_rename.ncu-rep gpurun_out/r2d_utf8_rename.txt > /dev/null 2>&1
rm -f gpurun_out/r2d_utf8_rename.ncu-rep
head -30 gpurun_out/r2d_utf8_rename.txt
