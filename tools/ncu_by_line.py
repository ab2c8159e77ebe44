#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source cuda,sass --csv` output per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv --kernel-id ::regex:NAME:1 | python tools/ncu_by_line.py [topN]"""
import csv
import sys
from collections import defaultdict

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file, hdr = None, None
agg = defaultdict(lambda: [0, 0, 0, 0, ""])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        wi = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else -1
        gi = hdr.index("L2 Theoretical Sectors Global")
        continue
    if hdr is None or len(r) != len(hdr) or r[0].strip() == "":
        continue  # rows without a line number are the SASS children of the preceding source row
    try:
        inst, samp, wf, gs = int(r[ii]), int(r[si]), (int(r[wi]) if wi >= 0 else 0), int(r[gi])
    except ValueError:
        continue
    a = agg[(cur_file, r[0])]
    a[0] += inst; a[1] += samp; a[2] += wf; a[3] += gs; a[4] = r[1].strip()[:100]
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
tw = sum(a[2] for a in agg.values()) or 1
print(f"total warp-instructions {ti}  samples {ts}  shared wavefronts {tw}")
print("--- by instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]/ti*100:5.1f}% inst {a[1]/ts*100:5.1f}% stall {a[2]/tw*100:5.1f}% smem  {k[0]}:{k[1]}  {a[4]}")
print("--- by stall samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
    print(f"{a[0]/ti*100:5.1f}% inst {a[1]/ts*100:5.1f}% stall {a[2]/tw*100:5.1f}% smem  {k[0]}:{k[1]}  {a[4]}")
