#!/usr/bin/env python
"""Small device-resident run of the hot path for ncu: `python tools/prof_run.py [ngal] [precision] [reps]`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import egg_b200  # noqa: E402
from egg_b200 import photometry  # noqa: E402
if os.environ.get('EGGPHOT_LIB'):
    photometry.load_library(os.environ['EGGPHOT_LIB'])
from egg_b200 import synth  # noqa: E402

np.seterr(all="ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32c"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
opt, ir = synth.make_opt_lib(), synth.ce01_lib()
gal = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5)
bset = [b for b in synth.B30 if b not in os.environ.get("EGGPHOT_DROP_BANDS", "").split(",")]
bands = synth.get_filters(bset)
ph = egg_b200.Photometry(opt, ir, bands, line_lambda=synth.LINE_LAMBDA, precision=prec,
                         smem_table_budget=int(os.environ.get("EGGPHOT_BUDGET", "0")))
cols = ["z", "d", "m_disk", "m_bulge", "m2lcor", "lir", "igm_abs", "vdisp", "line_lum_disk", "line_lum_bulge",
        "opt_sed_disk", "opt_sed_bulge", "ir_sed"]
dev = {k: torch.from_numpy(np.ascontiguousarray(gal[k]).view(np.int64) if gal[k].dtype == np.uint64
                           else np.ascontiguousarray(gal[k])).cuda() for k in cols}
outs = {k: torch.empty((n, len(bands)), dtype=torch.float32, device="cuda") for k in ("flux", "flux_disk", "flux_bulge")}
gp = {k: v.data_ptr() for k, v in dev.items()}
op = {k: v.data_ptr() for k, v in outs.items()}
st = torch.cuda.current_stream()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    ph.compute_device(gp, n, op, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    print(f"{prec} n={n}: {e0.elapsed_time(e1):.3f} ms -> {n*len(bands)/e0.elapsed_time(e1)*1e3:.3e} units/s", flush=True)
ph.synchronize()
print(ph.stats())
