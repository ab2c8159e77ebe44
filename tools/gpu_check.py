#!/usr/bin/env python
"""Quick on-GPU sanity run: parity statistics vs the oracle + a rough kernel timing."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import egg_b200  # noqa: E402
from egg_b200 import synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

np.seterr(all="ignore")


def relerr(a, b):
    floor = 1e-9 * np.abs(b).max(axis=1, keepdims=True)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(b), floor), 1e-300)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    opt, ir = synth.make_opt_lib(), synth.ce01_lib()
    gal = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5)
    bands, rf = synth.get_filters(synth.B30), synth.get_filters(synth.RF3)
    for kw in (dict(), dict(no_nebular=True)):
        ref = Oracle(opt, ir, bands, rf, line_lambda=synth.LINE_LAMBDA, **kw).compute(gal, f64=True)
        for prec in ("fp64", "fp32c"):
            ph = egg_b200.Photometry(opt, ir, bands, rf, line_lambda=synth.LINE_LAMBDA, precision=prec, **kw)
            t = time.time()
            out = ph.compute(gal, f64=True)
            dt = time.time() - t
            s = f"{kw} {prec}: "
            for k in ("flux_disk", "flux_bulge"):
                e = relerr(out[k + "_f64"], ref[k])
                s += f"{k} max {e.max():.2e} p99 {np.quantile(e, 0.99):.1e} | "
            for k in ("rfmag_disk", "rfmag_bulge"):
                s += f"{k} {np.abs(out[k + '_f64'] - ref[k]).max():.1e} | "
            print(s, f"{dt*1e3:.1f} ms", ph.stats(), flush=True)
            ph.close()
    # rough throughput: big batch, host path
    gal = synth.draw_galaxies(200000, opt_lib=opt, ir_lib=ir, mmin=7.5)
    for prec in ("fp32c", "fp64"):
        ph = egg_b200.Photometry(opt, ir, bands, line_lambda=synth.LINE_LAMBDA, precision=prec)
        ph.compute(gal)
        t = time.time()
        ph.compute(gal)
        dt = time.time() - t
        print(f"{prec}: 200k galaxies x 30 bands host path {dt*1e3:.1f} ms -> {200000*30/dt:.3e} units/s", flush=True)
        ph.close()


if __name__ == "__main__":
    main()
