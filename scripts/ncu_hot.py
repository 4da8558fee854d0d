#!/usr/bin/env python
"""Hot SASS regions of an `ncu --page source --csv` export: runs of instructions with similar execution counts."""
import csv, sys, gzip
f = gzip.open(sys.argv[1], 'rt') if sys.argv[1].endswith('.gz') else open(sys.argv[1])
r = csv.reader(f); next(r); h = next(r)
ie = h.index('Instructions Executed'); at = h.index('Avg. Threads Executed'); src = h.index('Source'); ws = h.index('Warp Stall Sampling (All Samples)')
rows = [x for x in r]
tot = sum(int(x[ie]) for x in rows); tots = sum(int(x[ws]) for x in rows)
print('total warp-instr', tot, 'static instr', len(rows), 'samples', tots)
cur = None; start = 0; acc = 0; accs = 0
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
for i, x in enumerate(rows + [None]):
    e = int(x[ie]) if x else -1
    if cur is None or x is None or abs(e - cur) > 0.2 * max(cur, 1):
        if cur is not None and (acc > thr * tot or accs > thr * tots):
            print('instr %4d-%4d  exec/instr ~%9d  threads %s  instr-share %.1f%%  stall-share %.1f%%' % (start, i - 1, cur, rows[start][at], 100 * acc / tot, 100 * accs / tots))
        cur = e; start = i; acc = 0; accs = 0
    if x: acc += e; accs += int(x[ws])
if len(sys.argv) > 3:
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][ws]))[:int(sys.argv[3])]
    for i in sorted(top):
        x = rows[i]; print('%4d %-64s samples %6d (%.1f%%) exec %8s thr %s' % (i, x[src].strip()[:64], int(x[ws]), 100 * int(x[ws]) / tots, x[ie], x[at]))
