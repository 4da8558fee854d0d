#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
   python scripts/launch_summary.py launches.csv [first_launch_to_count]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.OrderedDict()
for n, r in enumerate(rows[1:]):
    if n < skip: continue
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    if r[ui] == "ns": v /= 1e3
    elif r[ui] == "ms": v *= 1e3
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += v
s = sum(t[1] for t in tot.values())
print(f"{len(rows) - 1 - skip} launches, {s / 1e3:.2f} ms serialised")
for k, (n, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k:28s} n={n:4d} total {v / 1e3:9.3f} ms  avg {v / n:9.1f} us  {100 * v / s:5.1f} %")
