#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by source line:
share of executed warp instructions and of stall samples per line (developer tool)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, ""])
total = tsamp = 0


def num(s):
    try:
        return int(float(s.replace(",", "")))
    except ValueError:
        return 0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].strip().isdigit():
        continue
    d = dict(zip(hdr, r))
    inst, samp = num(d.get("Instructions Executed", "0")), num(d.get("# Samples", "0"))
    key = (cur_file, int(r[0]))
    agg[key][0] += inst
    agg[key][1] += samp
    agg[key][2] = r[1]
    total += inst
    tsamp += samp
print("total warp instructions", total, "samples", tsamp)
for (f, ln), (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} inst {inst / max(total, 1) * 100:5.1f}% samp {samp / max(tsamp, 1) * 100:5.1f}%  {src.strip()[:110]}")
