#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py list.csv"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    n = r[ki].split("(")[0][:70]; tot[n] += float(r[vi].replace(",", "")) * 1e-3; cnt[n] += 1
T = sum(tot.values())
for n in sorted(tot, key=lambda x: -tot[x]):
    print("%-72s launches %5d  total %11.1f us  avg %9.1f us  share %5.1f%%" % (n, cnt[n], tot[n], tot[n] / cnt[n], 100 * tot[n] / T))
