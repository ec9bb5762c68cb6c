"""Summarise `ncu --page source --csv --print-source=cuda` output: hottest source lines of a kernel.
usage: ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass > src.csv; python scripts/ncu_hot_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr, out, kern = None, None, {}, '?'
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if r and r[0] == 'Kernel Name':
        kern = r[1]
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        key = (kern, cur, int(r[0]))
        if key in out:
            continue
        out[key] = (r[1][:100], int(d.get('# Samples', 0) or 0), int(d['Instructions Executed']), int(d['Thread Instructions Executed']),
                    {k: int(v) for k, v in d.items() if k.startswith('stall_') and '(' not in k and v.isdigit() and int(v) > 0})
kernels = sorted(set(k[0] for k in out))
for kn in kernels[:1]:
    items = [(k, v) for k, v in out.items() if k[0] == kn]
    ts = sum(v[1] for _, v in items) or 1
    ti = sum(v[2] for _, v in items) or 1
    print(kn, 'samples', ts, 'warp-inst', ti)
    for k, v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
        st = sorted(v[4].items(), key=lambda kv: -kv[1])[:3]
        print('%-14s %4d smp=%5.1f%% inst=%5.1f%% thr/inst=%4.1f %-38s| %s' % (k[1], k[2], 100 * v[1] / ts, 100 * v[2] / ti, v[3] / max(v[2], 1),
              ','.join('%s:%d' % (a.replace('stall_', ''), b) for a, b in st), v[0]))

# ---- per-file / per-line-range phase totals (optional 3rd arg: file:lo-hi,file:lo-hi,...) ----
if len(sys.argv) > 3:
    for kn in kernels[:1]:
        items = [(k, v) for k, v in out.items() if k[0] == kn]
        ts = sum(v[1] for _, v in items) or 1
        ti = sum(v[2] for _, v in items) or 1
        for spec in sys.argv[3].split(','):
            f, rng = spec.split(':')
            lo, hi = map(int, rng.split('-'))
            sel = [v for k, v in items if k[1] == f and lo <= k[2] <= hi]
            st = {}
            for v in sel:
                for a, b in v[4].items():
                    st[a] = st.get(a, 0) + b
            top3 = sorted(st.items(), key=lambda kv: -kv[1])[:4]
            print('%-28s smp=%5.1f%% inst=%5.1f%%  %s' % (spec, 100 * sum(v[1] for v in sel) / ts, 100 * sum(v[2] for v in sel) / ti,
                  ','.join('%s:%d' % (a.replace('stall_', ''), b) for a, b in top3)))
