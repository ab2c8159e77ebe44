#!/usr/bin/env python
"""Parity report: the CUDA path (through the C ABI) against the oracle on every case of tests/cases.py at a larger
sample than the unit tests, both arithmetic modes.  Writes error statistics as JSON (committed under profiles/).
usage: python tools/parity_report.py [ngal] > profiles/r1_parity_report.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import egg_b200  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

np.seterr(all="ignore")
ngal = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
report = {"ngal_per_case": ngal, "definition": "rel = |cuda - oracle| / |oracle| per (galaxy, band) entry, restricted to entries brighter "
          "than `floor` x the galaxy's brightest band; rfmag: absolute difference in mag", "cases": {}}
for name in sorted(cases.CASES):
    c = cases.make(name, ngal=ngal, seed=77)
    ref = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"]).compute(c["gal"], f64=True)
    entry = {"nband": len(c["bands"]), "nrf": len(c["rfbands"]), "options": {k: v for k, v in c["opts"].items()}}
    for prec in ("fp32c", "fp64"):
        ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], precision=prec, **c["opts"])
        out = ph.compute(c["gal"], f64=True)
        ph.close()
        e = {}
        for k in ("flux_disk", "flux_bulge"):
            a, b = out[k + "_f64"], ref[k]
            bright = np.abs(b).max(axis=1, keepdims=True)
            for floor in (1e-5, 1e-9, 0.0):
                m = (np.abs(b) > floor * bright) & (b != 0)
                r = np.abs(a - b)[m] / np.abs(b)[m]
                e[f"{k}_floor{floor:g}"] = {"max": float(r.max()) if r.size else 0.0,
                                            "p99.9": float(np.quantile(r, 0.999)) if r.size else 0.0,
                                            "median": float(np.median(r)) if r.size else 0.0, "entries": int(m.sum())}
            # entries the oracle has exactly 0 (uncovered band -> 0, or every SED node under the band erased by the
            # IGM): the largest value the CUDA path has there, relative to the galaxy's brightest band
            z = b == 0
            e[f"{k}_max_where_oracle_is_zero_rel_brightest"] = float((np.abs(a) / np.maximum(bright, 1e-300))[z].max()) if z.any() else 0.0
        for k in ("rfmag_disk", "rfmag_bulge"):
            if ref[k].size:
                e[k + "_max_abs"] = float(np.abs(out[k + "_f64"] - ref[k]).max())
        entry[prec] = e
    report["cases"][name] = entry
    print(name, "done", file=sys.stderr, flush=True)
json.dump(report, sys.stdout, indent=1)
