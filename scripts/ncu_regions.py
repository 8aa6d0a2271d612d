"""Summarise an `ncu --page source --print-source sass --csv` dump: instructions executed and stall samples per code
region, where regions are given as SASS offset ranges (hex, relative to the kernel start) on the command line:
    python scripts/ncu_regions.py file.csv name:lo-hi name:lo-hi ...   (no ranges: print the top lines by samples)"""
import csv, sys

def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    out = []
    base = None
    for r in rows[2:]:
        if len(r) < len(hdr) - 5:
            continue
        d = dict(zip(hdr, r))
        addr = int(d['Address'], 16)
        base = addr if base is None else base
        d['_off'] = addr - base
        out.append(d)
    return hdr, out

def main():
    hdr, rows = load(sys.argv[1])
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    total_inst = sum(int(r['Instructions Executed']) for r in rows)
    total_samp = sum(int(r['# Samples']) for r in rows)
    print(f"total warp-instructions {total_inst}, samples {total_samp}")
    regions = []
    for a in sys.argv[2:]:
        name, rng = a.split(':')
        lo, hi = (int(x, 16) for x in rng.split('-'))
        regions.append((name, lo, hi))
    if not regions:
        top = sorted(rows, key=lambda r: -int(r['# Samples']))[:40]
        for r in top:
            st = {c: int(r[c]) for c in stall_cols if int(r[c])}
            print(f"{r['_off']:#06x} {r['Source'].strip():<60} inst {r['Instructions Executed']:>9} samples {r['# Samples']:>6} {st}")
        return
    for name, lo, hi in regions:
        sel = [r for r in rows if lo <= r['_off'] <= hi]
        inst = sum(int(r['Instructions Executed']) for r in sel)
        samp = sum(int(r['# Samples']) for r in sel)
        st = {c[6:]: sum(int(r[c]) for r in sel) for c in stall_cols}
        st = {k: v for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > samp * 0.02}
        print(f"{name:<14} {lo:#06x}-{hi:#06x}: inst {inst:>10} ({100*inst/total_inst:5.1f} %)  samples {samp:>7} ({100*samp/total_samp:5.1f} %)  {st}")

if __name__ == '__main__':
    main()
