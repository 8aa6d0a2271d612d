"""Per-source-line summary of an `ncu --page source --print-source cuda,sass --csv` dump: instructions executed, samples
and the top stall reasons of every CUDA-C line that matters (>= 1 % of samples or instructions)."""
import csv, sys, collections

def main(path, min_pct=1.0):
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        extra = len(r) - len(hdr)  # unescaped quotes/commas in the source text split it into more fields
        r = [r[0], ','.join(r[1:2 + extra])] + r[2 + extra:]
        d = dict(zip(hdr[4:], r[4:]))
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, {'src': r[1].strip(), 'inst': 0, 'samp': 0, 'st': collections.Counter()})
        a['inst'] += int(d['Instructions Executed'])
        a['samp'] += int(d['# Samples'])
        for k, v in d.items():
            if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v):
                a['st'][k[6:]] += int(v)
    ti = sum(a['inst'] for a in agg.values()); ts = sum(a['samp'] for a in agg.values())
    print(f"total warp-instructions {ti}, samples {ts}")
    for (f, ln), a in agg.items():
        if a['inst'] * 100 >= ti * min_pct or a['samp'] * 100 >= ts * min_pct:
            st = ' '.join(f"{k}:{v}" for k, v in a['st'].most_common(4))
            print(f"{f}:{ln:<4} inst {100*a['inst']/ti:5.1f}% samp {100*a['samp']/ts:5.1f}%  {st:<60} | {a['src'][:90]}")

if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
