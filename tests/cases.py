"""Shared parity cases: (name, libs, bands, rfbands, options) used by the CPU-tier emulation tests, the
golden fixtures and the GPU parity tests, so that all three exercise the same inputs."""
import numpy as np

from egg_b200 import synth

_cache = {}


def libs():
    if not _cache:
        _cache["opt"] = synth.make_opt_lib()
        _cache["opt_hd"] = synth.make_opt_lib(seed=3, hd=True)
        _cache["ce01"] = synth.ce01_lib()
        _cache["cs17"] = synth.make_cs17_lib()
    return _cache


# name -> (opt key, ir key, bands, rfbands, opts)
CASES = {
    "b4_defaults": ("opt", "ce01", synth.B4, [], {}),
    "b30_rf3_defaults": ("opt", "ce01", synth.B30, synth.RF3, {}),
    "b30_no_nebular": ("opt", "ce01", synth.B30, synth.RF3, {"no_nebular": True}),
    "b40_hd_cs17": ("opt_hd", "cs17", synth.B40, synth.RF3, {}),
    "b30_igm_none": ("opt", "ce01", synth.B30, [], {"igm": "none"}),
    # BASELINE.json configs[4]: the high-definition library without IGM (opt_lib_fast_hd_noigm), run in fp64 with save_sed
    "b30_hd_noigm": ("opt_hd", "ce01", synth.B30, synth.RF3, {"igm": "none"}),
    "b30_line_average": ("opt", "ce01", synth.B30, [], {"line_profile": "average"}),
    "b23_no_dust": ("opt", "ce01", synth.B23, synth.RF3, {"no_dust": True}),
    "b23_no_stellar": ("opt", "ce01", synth.B23, synth.RF3, {"no_stellar": True, "igm": "none"}),
    "b4_chabrier_flags": ("opt", "ce01", synth.B4 + ["spitzer-irac1", "herschel-pacs100"], [],
                          {"opt_lib_imf": "chabrier", "trim_filters": True, "filter_photons": True}),
    "b4_flambda_nonorm": ("opt", "ce01", synth.B4, [], {"filter_flambda": True, "filter_normalize": False,
                                                        "opt_lib_min_step": 0.0}),
    # the reference's bulges carry no lines (src/egg-gencat.cpp:1886-1976) but get_flux would add them: exercise it
    "b23_bulge_lines": ("opt", "ce01", synth.B23, synth.RF3, {"_bulge_lines": 0.3}),
}


def make(name, ngal=96, seed=11):
    ok, ik, bands, rf, opts = CASES[name]
    L = libs()
    gal = synth.draw_galaxies(ngal, opt_lib=L[ok], ir_lib=L[ik], mmin=7.5, seed=seed)
    opts = dict(opts)
    frac = opts.pop("_bulge_lines", None)
    if frac is not None:
        gal["line_lum_bulge"] = (frac * gal["line_lum_disk"]).astype(np.float32)
    return dict(opt=L[ok], ir=L[ik], bands=synth.get_filters(bands), rfbands=synth.get_filters(rf),
                opts=dict(opts), gal=gal, lines=synth.LINE_LAMBDA)


FLOOR = {"fp32c": 0.0, "fp64": 0.0}   # no faint-band floor: the tolerance holds relative to every band itself


def rel_err(a, b, floor_frac=0.0):
    """|a-b| / |b| per entry (BASELINE.json north_star: relative, per band).  Where the reference value is exactly 0
    (uncovered band, or every SED node under the band erased by the IGM) the result must be exactly 0 too: anything
    else counts as an infinite error.  floor_frac > 0 (not used by the parity tests any more) measures faint entries
    against that fraction of the galaxy's brightest band instead."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if b.size == 0:
        return np.zeros_like(b)
    den = np.abs(b)
    if floor_frac > 0:
        den = np.maximum(den, floor_frac * np.abs(b).max(axis=1, keepdims=True))
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.abs(a - b) / den
    e[(den == 0) & (a == b)] = 0.0
    e[(den == 0) & (a != b)] = np.inf
    return e


TOL = {"fp32c": 1e-5, "fp64": 1e-10}   # BASELINE.json north_star tolerances (relative, per band)
MAGTOL = {"fp32c": 2e-5, "fp64": 1e-9}  # magnitudes: 2.5 log10(e) x the flux tolerance, absolute
