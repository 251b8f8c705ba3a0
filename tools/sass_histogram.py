#!/usr/bin/env python
"""SASS opcode histogram per kernel of the shipped library (cuobjdump -sass | c++filt).  Usage: python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "qoc_b200", "libqocb200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
names = [p.split("\n", 1)[0].strip() for p in parts[1:]]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
print("# SASS opcode histogram per kernel of qoc_b200/libqocb200.so (cuobjdump -sass, sm_100a), round 2 final build")
print("# DMMA.8x8x4 = FP64 tensor MMA; UTMALDG = cp.async.bulk.tensor (TMA); SYNCS.* = mbarrier ops; LDGSTS = cp.async")
print()
KEEP = ("DMMA", "UTMALDG", "SYNCS", "LDGSTS", "FENCE", "UBLKCP", "REDUX", "BAR")
tot = collections.Counter()
for p, name in zip(parts[1:], dem):
    c = collections.Counter()
    n = 0
    for l in p.split("\n"):
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", l)
        if not m or "/*" not in l.split(m.group(1))[0]:
            continue
        op = m.group(1)
        n += 1
        key = op if op.startswith(KEEP) else op.split(".")[0]
        if key.startswith("DMMA"):
            key = "DMMA.8x8x4"
        c[key] += 1
        tot[key] += 1
    top = [k for k, _ in c.most_common(10)] + [k for k in c if k.startswith(("UTMALDG", "SYNCS", "LDGSTS", "FENCE", "UBLKCP"))]
    seen, items = set(), []
    for k in top:
        if k not in seen:
            seen.add(k); items.append("%s=%d" % (k, c[k]))
    print(name[:160])
    print("    instructions %d : %s" % (n, "  ".join(items)))
print()
print("# whole library: " + "  ".join("%s=%d" % (k, tot[k]) for k in ("DMMA.8x8x4",) + tuple(sorted(k for k in tot if k.startswith(("UTMALDG", "SYNCS", "LDGSTS", "UBLKCP"))))))
