#!/usr/bin/env python
"""Large-sample parity sweep of one case: `python tools/parity_sweep.py <case> <ngal> [seed]` (GPU + oracle)."""
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
name, ngal = sys.argv[1], int(sys.argv[2])
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 123
c = cases.make(name, ngal=ngal, seed=seed)
ref = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"]).compute(c["gal"], f64=True)
for prec in ("fp32c", "fp64"):
    ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], precision=prec, **c["opts"])
    out = ph.compute(c["gal"], f64=True)
    ph.close()
    for k in ("flux_disk", "flux_bulge"):
        e = cases.rel_err(out[k + "_f64"], ref[k])  # relative to the band itself, no floor
        w = np.unravel_index(np.argmax(e), e.shape)
        nz = int(((ref[k] == 0) & (out[k + "_f64"] != 0)).sum())
        print(f"{name} n={ngal} {prec} {k}: max rel (no floor) {e.max():.2e} (galaxy {w[0]}, band {w[1]}, z={c['gal']['z'][w[0]]:.4f}); "
              f"entries above {cases.TOL[prec]:g}: {(e > cases.TOL[prec]).sum()} of {e.size}; non-zero where the oracle has 0: {nz}", flush=True)
    for k in ("rfmag_disk", "rfmag_bulge"):
        if ref[k].size:
            print(f"   {k} max abs {np.abs(out[k + '_f64'] - ref[k]).max():.2e}")
