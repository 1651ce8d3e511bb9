#!/usr/bin/env python
"""Mnemonic counts per kernel of the built library (evidence that the hot kernels are Blackwell-native):
  python scripts/sass_counts.py > profiles/sass_r2.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tools_b200", "libqfall_b200.so")
COLS = ["UTCIMMA", "UTCBAR", "UTCBAR.MULTICAST", "LDTM", "STTM", "UTMALDG", "UTMALDG.3D.MULTICAST", "LDGSTS", "DMMA", "DFMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
mangled = sorted(set(re.findall(r"Function : (\S+)", sass)))
dem = dict(zip(mangled, subprocess.run(["cu++filt"] + mangled, capture_output=True, text=True).stdout.splitlines())) if mangled else {}
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = dem.get(m.group(1), m.group(1))
        name = re.sub(r"\(.*", "", name).replace("(anonymous namespace)::", "").replace("void ", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    for c in COLS:
        if op == c or (op.startswith(c + ".") and not any(op.startswith(o) for o in COLS if o != c and o.startswith(c) and len(o) > len(c))):
            cur[c] += 1
    if op.startswith("UTMALDG") and "MULTICAST" in op:
        cur["UTMALDG.3D.MULTICAST"] += 0  # counted above through startswith
print("# SASS evidence (cuobjdump -sass tools_b200/libqfall_b200.so, sm_100a), round 2\n")
print("Mnemonic counts per kernel (`tcgen05.mma` -> UTCIMMA, `tcgen05.ld/st` -> LDTM/STTM, TMA -> UTMALDG [multicast: "
      "UTMALDG.3D.MULTICAST],\n`tcgen05.commit` -> UTCBAR [onto both CTAs of a pair: UTCBAR.MULTICAST], `cp.async` -> LDGSTS, "
      "`mma.sync` f64 -> DMMA).\nRegenerate with `python scripts/sass_counts.py > profiles/sass_r2.md`.\n")
print("| kernel (all template instantiations) | " + " | ".join(COLS) + " |\n|---|" + "---|" * len(COLS))
tot = collections.Counter()
for name, c in counts.items():
    if not any(c[k] for k in COLS if k != "DFMA"):
        continue
    print(f"| `{name}` | " + " | ".join(str(c[k]) for k in COLS) + " |")
    tot.update(c)
print("| **library total (kernels above)** | " + " | ".join(str(tot[k]) for k in COLS) + " |")
