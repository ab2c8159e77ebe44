#!/usr/bin/env python
"""Generate tests/golden/*.npz: oracle outputs (double, unrounded) for fixed seeded inputs.

The reference cannot be built or imported here (its numerics live in the absent vif library), so
these vectors pin the ORACLE restatement against regressions; they are not outputs of egg itself.
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

np.seterr(all="ignore")
GOLDEN = ("b4_defaults", "b30_rf3_defaults", "b40_hd_cs17", "b23_no_dust", "b30_hd_noigm")
for name in (sys.argv[1:] or GOLDEN):  # (names on the command line: only those files are rewritten)
    c = cases.make(name, ngal=256, seed=2024)
    ref = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"]).compute(c["gal"], f64=True)
    out = {k: ref[k] for k in ("flux_disk", "flux_bulge", "rfmag_disk", "rfmag_bulge")}
    # a fingerprint of the inputs, so that a change of the synthetic generators is noticed
    out["input_checksum"] = np.array([np.float64(c["gal"]["z"].sum()), np.float64(c["gal"]["lir"].astype(np.float64).sum()),
                                      np.float64(c["opt"]["sed"].astype(np.float64).sum())])
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))
