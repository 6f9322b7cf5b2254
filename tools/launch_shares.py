#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mu = hdr.index('Metric Unit')
t = defaultdict(float); c = defaultdict(int)
for r in rows[h + 1:]:
    if len(r) > mv:
        v = float(r[mv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[mu], 1e-3)
        t[r[kn]] += v; c[r[kn]] += 1
tot = sum(t.values())
print(f"total {tot/1e3:.3f} ms over {sum(c.values())} launches")
for k, v in sorted(t.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v:12.1f} us {c[k]:5d} {100*v/tot:5.1f}%  {k[:100]}")
