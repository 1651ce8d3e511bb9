#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list, restricted to the launches after
`PROFILE_REGION_BEGIN launches so far: N` (scripts/prof_step.py):  python scripts/launch_summary.py launches.csv prof.log"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 2:]
    idx = {h: i for i, h in enumerate(hdr)}
    start = 0
    if len(sys.argv) > 2:
        m = re.search(r"launches so far: (\d+)", open(sys.argv[2]).read())
        start = int(m.group(1)) if m else 0
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if int(r[idx["ID"]]) < start:
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1}.get(r[idx["Metric Unit"]], 1e-6)
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[:90]} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.2f} | |")


if __name__ == "__main__":
    main()
