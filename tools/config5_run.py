#!/usr/bin/env python
"""BASELINE.json configs[4] through the command line, timed end to end: egg-gencat-b200 input_props=... save_sed precision=fp64
with the high-definition optical library without IGM (opt_lib_fast_hd_noigm) on N galaxies (default 1.95e5), 30 bands.
Prints the wall time of the program, the size of the spectra file and the rate at which it was produced.
usage: python tools/config5_run.py [ngal] [workdir]"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from egg_b200 import fitscol, synth  # noqa: E402
from test_cli import CLI, _filter_db  # noqa: E402
import pathlib  # noqa: E402

np.seterr(all="ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 195_000
work = pathlib.Path(sys.argv[2] if len(sys.argv) > 2 else "/tmp/egg_config5")
work.mkdir(parents=True, exist_ok=True)
L = cases.libs()
opt, ir = L["opt_hd"], L["ce01"]
gal = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5, seed=5)
bands = synth.B30
db = _filter_db(work, bands)
fitscol.write_table(str(work / "opt_lib.fits"), {k.upper(): np.asarray(v) for k, v in opt.items()})
fitscol.write_table(str(work / "ir_lib_ce01.fits"), {"LAM": np.asarray(ir["lam"]), "SED": np.asarray(ir["sed"])})
cat = {"ID": np.arange(n, dtype=np.uint64), "RA": np.zeros(n), "DEC": np.zeros(n)}
for k in ("z", "d", "m_disk", "m_bulge", "m2lcor", "lir", "igm_abs", "vdisp", "line_lum_disk", "line_lum_bulge", "opt_sed_disk", "opt_sed_bulge", "ir_sed"):
    cat[k.upper()] = np.asarray(gal[k])
cat["LINES"] = np.array(synth.LINES)
cat["LINE_LAMBDA"] = synth.LINE_LAMBDA
fitscol.write_table(str(work / "cat.fits"), cat)
args = [CLI, f"input_props={work}/cat.fits", f"filter_db={db}", f"opt_lib={work}/opt_lib.fits", f"ir_lib={work}/ir_lib_ce01.fits",
        "bands=[" + ",".join(bands) + "]", "igm=none", "precision=fp64", f"out={work}/out.fits"]
for extra, tag in (([], "fluxes only"), (["save_sed"], "fluxes + save_sed")):
    t0 = time.perf_counter()
    r = subprocess.run(args + extra, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stderr)
        raise SystemExit(1)
    line = f"{tag}: {n} galaxies x {len(bands)} bands, fp64, hd library without IGM: {dt:.2f} s wall"
    sf = work / "out-seds.dat"
    if extra and sf.exists():
        gb = sf.stat().st_size / 1e9
        line += f"; spectra file {gb:.2f} GB"
        t_sed = dt - t_flux
        line += f"; save_sed adds {t_sed:.2f} s = {gb / max(t_sed, 1e-9):.2f} GB/s from kernel to file (PCIe gen5 x16 ~ 55 GB/s one way; the file system of the box is the slower end)"
    else:
        t_flux = dt
    print(line, flush=True)
for f in ("out-seds.dat", "out.fits", "cat.fits"):
    try:
        os.remove(work / f)
    except OSError:
        pass
