#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line / 25-line bucket."""
import csv, sys, gzip, collections
f = gzip.open(sys.argv[1], 'rt') if sys.argv[1].endswith('.gz') else open(sys.argv[1])
r = csv.reader(f)
cur = None; agg = collections.Counter(); aggi = collections.Counter(); srcs = {}
def num(x):
    try: return int(x)
    except Exception: return 0
for row in r:
    if len(row) == 2 and row[0] == 'File Path': cur = row[1].split('/')[-1]; continue
    if len(row) < 8 or row[0] in ('Line No', 'Function Name', ''): continue
    try: ln = int(row[0])
    except Exception: continue
    agg[(cur, ln)] += num(row[4]); aggi[(cur, ln)] += num(row[7]); srcs[(cur, ln)] = row[1].strip()[:90]
tot = sum(agg.values()) or 1; ti = sum(aggi.values()) or 1
print('total samples', tot, 'warp instr', ti)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 25
b = collections.Counter(); bi = collections.Counter()
for (fn, l), v in agg.items(): b[(fn, l // B * B)] += v
for (fn, l), v in aggi.items(): bi[(fn, l // B * B)] += v
for k, v in b.most_common(40): print('%-16s %4d  samples %5.1f%%  instr %5.1f%%' % (k[0], k[1], 100 * v / tot, 100 * bi[k] / ti))
print()
for k, v in agg.most_common(30): print('%-16s %4d %5.1f%% %5.1f%%  %s' % (k[0], k[1], 100 * v / tot, 100 * aggi[k] / ti, srcs[k]))
