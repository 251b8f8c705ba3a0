#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean duration and
share of the total.  Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[ki][:70], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("# source: %s  (cold-cache, serialised: compare SHARES, not absolutes)" % sys.argv[1])
print("%-72s %-14s %-14s %5s %12s %7s" % ("kernel", "grid", "block", "n", "avg_us", "share"))
for (k, g, b), v in agg.items():
    print("%-72s %-14s %-14s %5d %12.1f %7.3f" % (k, g, b, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
