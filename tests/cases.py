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


FLOOR = {"fp32c": 1e-5, "fp64": 1e-9}


def rel_err(a, b, floor_frac=1e-9):
    """|a-b| / max(|b|, floor) per entry; floor = floor_frac x the galaxy's brightest band.  A band that
    far below the brightest one (an IGM-erased band seen only through the 1e-6 wing of its filter) is
    held to the tolerance in absolute terms relative to that floor, not relative to itself.  Measured
    worst cases relative to the band itself: fp32c 4e-5 for a band 4e-7 of the brightest whose flux is a
    strong emission line seen through the 1e-6 wing of an HST filter (the hi+lo prefix tables hold
    2^-48 of the band total), < 3e-6 for every band above 1e-6 of the brightest; fp64 2e-12."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if b.size == 0:
        return np.zeros_like(b)
    floor = floor_frac * np.abs(b).max(axis=1, keepdims=True)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(b), floor), 1e-300)


TOL = {"fp32c": 1e-5, "fp64": 1e-10}   # BASELINE.json north_star tolerances (relative, per band)
MAGTOL = {"fp32c": 2e-5, "fp64": 1e-9}  # magnitudes: 2.5 log10(e) x the flux tolerance, absolute
