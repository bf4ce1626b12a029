"""Bin the warp-state samples of an `ncu --page source --csv` export by runs of SASS instructions (development aid).
  python tools/sass_samples.py source.csv [bin]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
h, data = rows[1], rows[2:]
si, src, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'instrs', len(data))
for b in range(0, len(data), B):
    chunk = data[b:b + B]
    s = sum(int(r[si]) for r in chunk)
    ops = {}
    for r in chunk:
        t = r[src].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    ex = sum(int(r[ie]) for r in chunk)
    hot = max(chunk, key=lambda r: int(r[si]))
    print('%5d-%5d samples %6d (%4.1f%%) exec %6d  %s  hot: %s (%s)' % (b, b + B, s, 100 * s / max(tot, 1), ex, top, hot[src].strip()[:50], hot[si]))
