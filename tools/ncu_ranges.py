#!/usr/bin/env python
"""Instruction and stall-sample shares per source-line range:
ncu -i rep --page source --print-source cuda,sass --csv | python tools/ncu_ranges.py file.cu a-b:name ..."""
import csv
import sys
from collections import defaultdict

fname = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    r, name = a.split(":")
    lo, hi = r.split("-")
    ranges.append((int(lo), int(hi), name))
rows = list(csv.reader(sys.stdin))
cur, hdr = None, None
inst, samp = defaultdict(int), defaultdict(int)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) != len(hdr) or r[0].strip() == "":
        continue
    try:
        inst[(cur, int(r[0]))] += int(r[ii])
        samp[(cur, int(r[0]))] += int(r[si])
    except ValueError:
        pass
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
byf = defaultdict(lambda: [0, 0])
for (f, l), v in inst.items():
    byf[f][0] += v
for (f, l), v in samp.items():
    byf[f][1] += v
print(f"total warp-instructions {ti}, samples {ts}")
for f, (a, b) in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:32s} {a/ti*100:5.1f}% inst {b/ts*100:5.1f}% samples")
for lo, hi, name in ranges:
    a = sum(v for (f, l), v in inst.items() if f == fname and lo <= l <= hi)
    b = sum(v for (f, l), v in samp.items() if f == fname and lo <= l <= hi)
    print(f"  {name:24s} {lo:4d}-{hi:4d} {a/ti*100:5.1f}% inst {b/ts*100:5.1f}% samples")
