#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line.
Usage: python tools/ncu_hot_lines.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
fname = None
data = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        si = [i for i, h in enumerate(hdr) if h == "# Samples"][0]
        ii = [i for i, h in enumerate(hdr) if h == "Instructions Executed"][0]
        continue
    if hdr and len(r) > 6 and r[0] not in ("", "Line No"):
        try:
            data.append((fname, int(r[0]), r[1].strip(), int(r[si]), int(r[ii])))
        except ValueError:
            pass
tot = sum(d[3] for d in data)
print("total samples", tot)
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%-16s %5d %8d %5.1f%% inst=%-10d %s" % (d[0], d[1], d[3], 100.0 * d[3] / max(tot, 1), d[4], d[2][:100]))
