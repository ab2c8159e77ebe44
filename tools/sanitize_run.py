#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
`compute-sanitizer --tool memcheck python tools/sanitize_run.py`.  4096 galaxies through the flux kernel in both
arithmetic modes (30 bands + 3 rest-frame bands, lines on) and through the property kernel."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import egg_b200  # noqa: E402
from egg_b200 import properties, synth  # noqa: E402

np.seterr(all="ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
opt, ir = synth.make_opt_lib(), synth.ce01_lib()
gal = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5, seed=5)
bands, rf = synth.get_filters(synth.B30), synth.get_filters(synth.RF3)
for prec in ("fp32c", "fp64"):
    ph = egg_b200.Photometry(opt, ir, bands, rf, line_lambda=synth.LINE_LAMBDA, precision=prec)
    out = ph.compute(gal)
    print(prec, "flux checksum", float(out["flux"].astype(np.float64).sum()), "launches", ph.stats()["kernel_launches"], flush=True)
    ph.close()
rng = np.random.default_rng(0)
z, m = rng.uniform(0.1, 8, n).astype(np.float32), rng.uniform(8, 11.5, n).astype(np.float32)
pr = properties.Properties(opt, seed=3)
got = pr.compute(z, m, rng.random(n) < 0.25)
pr.close()
print("props checksum", float(got["lir"].astype(np.float64).sum()), flush=True)
