#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares, not absolutes)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0][:70]; v = float(r[vi].replace(',', ''))
    v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}.get(r[ui], v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'n':>5s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:5d} {t / 1e3:10.3f} {t / n:10.1f} {100 * t / tot:6.1f}%")
