"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (development aid)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split('(')[0]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(',', ''))
for n, (c, t) in agg.items():
    print('%-44s %4d launches %10.1f us total %8.1f us avg' % (n, c, t / 1e3, t / 1e3 / c))
