#!/usr/bin/env python
"""Static SASS instruction counts per kernel of the built library: `python tools/sass_summary.py > profiles/<tag>_sass_summary.txt`.
UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier operations (the Blackwell data-movement path of the walk kernel)."""
import collections
import os
import re
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "egg_b200", "libeggphot.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
ops = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "DSETP", "FFMA", "FMUL", "FADD", "F2F", "I2F", "F2I", "MUFU", "SHFL", "LDS", "STS",
       "LDG", "STG", "LDL", "STL", "ATOMG", "RED", "BAR", "WARPSYNC", "VOTE", "UTCHMMA", "HMMA"]
print("# SASS instruction counts per kernel of egg_b200/libeggphot.so (cuobjdump -sass, sm_100a; static counts, not executed counts)")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier operations.  No UTC*MMA / HMMA: the path is a quadrature, not a contraction.")
print("%-44s %6s " % ("kernel", "total") + " ".join("%6s" % o[:6] for o in ops))
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    m = re.search(r"\d+(lane_\w+?kernel|flux_kernel|props_kernel|sed_kernel|zsort_[a-z_]+|fma_peak_kernel)(I[a-zA-Z0-9]+?E)?(?:E|v|P)", name)
    label = (m.group(1) + (("<" + m.group(2)[1:-1] + ">") if m.group(2) else "")) if m else name[:42]
    c = collections.Counter()
    for l in f.split("\n"):
        mm = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if mm:
            c[mm.group(1)] += 1
    print("%-44s %6d " % (label[:44], sum(c.values())) + " ".join("%6d" % sum(v for k, v in c.items() if k.startswith(o)) for o in ops))
