#!/usr/bin/env python
"""Kernel time shares from an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki][:75]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 10]:
    print(f"{k:78s} n={v[0]:3d} total={v[1]/1e6:9.3f} ms share={v[1]/tot:.4f}")
