set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2g_launches.csv python scripts/utf8_probe.py 100000 > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2g_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:70]; agg.setdefault(k, []).append(float(r[vi].replace(",", ""))/1000)
for k, v in agg.items(): print(f"{len(v):4d} launches, mean {sum(v)/len(v):9.1f} us  min {min(v):8.1f} max {max(v):8.1f}  {k}")
PY
