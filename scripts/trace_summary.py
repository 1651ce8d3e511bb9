#!/usr/bin/env python
"""Summarise a QF_TRACE=1 log (per-launch lines of the int8 contraction kernel) by launch class.
Usage: python scripts/trace_summary.py gpurun_out/trace.log [min_B]"""
import collections
import re
import sys

min_b = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
rows = []
for l in open(sys.argv[1]):
    m = re.search(r"\[i8\] B=(\d+) N=(\d+) K=(\d+) LX=(\d+) LW=(\d+) kind=(\d+)\s+([\d.]+) ms\s+pairs/mac=([\d.]+)", l)
    if not m:
        continue
    B, N, K, LX, LW, kind = map(int, m.groups()[:6])
    ms, pm = float(m.group(7)), float(m.group(8))
    if B < min_b:
        continue
    rows.append((N, K, LX, LW, kind, ms, pm))
    cls = "K<=320" if K <= 320 else "K~1024" if K <= 1100 else "K>=2048"
    a = agg[(kind, cls, LW)]
    a[0] += 1
    a[1] += ms
    a[2] += 2.0 * B * N * K * pm
tot = sum(a[1] for a in agg.values())
print("| out_kind | K class | LW | launches | ms | share | TOP/s executed |\n|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {k[0]} | {k[1]} | {k[2]} | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f}% | {a[2] / a[1] / 1e9:.0f} |")
print(f"total {tot:.2f} ms")
print("launches above 0.5 ms: (N, K, LX, LW, kind, ms, pairs/mac)")
for r in rows:
    if r[5] > 0.5:
        print(" ", r)
