import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import egg_b200
from egg_b200 import photometry
if os.environ.get('EGGPHOT_LIB'): photometry.load_library(os.environ['EGGPHOT_LIB'])
from egg_b200 import synth
from oracle.oracle import Oracle
np.seterr(all="ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "fp64"
bset = getattr(synth, sys.argv[3]) if len(sys.argv) > 3 else synth.B4
neb = len(sys.argv) > 4 and sys.argv[4] == "lines"
opt, ir = synth.make_opt_lib(), synth.ce01_lib()
gal = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5)
bands = synth.get_filters(bset)
ref = Oracle(opt, ir, bands, [], line_lambda=synth.LINE_LAMBDA, no_nebular=not neb).compute(gal, f64=True)
ph = egg_b200.Photometry(opt, ir, bands, [], line_lambda=synth.LINE_LAMBDA, precision=prec, no_nebular=not neb)
out = ph.compute(gal, f64=True)
for k in ("flux_disk", "flux_bulge"):
    a, b = out[k + "_f64"], ref[k]
    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-9 * np.abs(b).max(axis=1, keepdims=True))
    print(sys.argv[1:], k, "max", e.max(), "bad bands", np.unique(np.argwhere(e > 1e-6)[:, 1]), "bad gals", np.unique(np.argwhere(e > 1e-6)[:, 0])[:20])
print(ph.stats())
