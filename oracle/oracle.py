"""ctypes driver for the CPU oracle (oracle/eggphot_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (egg_b200/) never imports this.

``Oracle`` mirrors the part of ``vif_main`` that feeds ``get_flux`` (src/egg-gencat.cpp:316-378
filters, :456-534 optical library, :2024-2240 photometry) and returns the same column groups.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libeggphot_oracle.so")

NODES = {"merged": 0, "sed": 1, "filter": 2}
GAUSS = {"integral": 0, "average": 1}

DEFAULT_OPTS = dict(igm="inoue14", no_stellar=False, no_dust=False, no_nebular=False,
                    opt_lib_imf="salpeter", opt_lib_min_step=0.001,
                    filter_flambda=False, filter_photons=False, filter_normalize=True, trim_filters=False,
                    sed2flux_nodes="merged", line_profile="integral", emulate_f32=False)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "eggphot_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


class _Setup(C.Structure):
    _fields_ = [("nopt_sed", C.c_int64), ("nopt", C.c_int64), ("opt_lam", C.c_void_p), ("opt_sed", C.c_void_p),
                ("ir_kind", C.c_int32), ("nir_sed", C.c_int64), ("nir", C.c_int64),
                ("ir_lam", C.c_void_p), ("ir_sed", C.c_void_p), ("ir_pah", C.c_void_p),
                ("nband", C.c_int32), ("foff", C.c_void_p), ("flam", C.c_void_p), ("fres", C.c_void_p),
                ("nrf", C.c_int32), ("rfoff", C.c_void_p), ("rflam", C.c_void_p), ("rfres", C.c_void_p),
                ("nline", C.c_int32), ("line_lambda", C.c_void_p),
                ("apply_igm", C.c_int32), ("no_stellar", C.c_int32), ("no_nebular", C.c_int32),
                ("nodes_mode", C.c_int32), ("gauss_mode", C.c_int32), ("emulate_f32", C.c_int32)]


class _Gal(C.Structure):
    _fields_ = [("ngal", C.c_int64), ("z", C.c_void_p), ("d", C.c_void_p), ("m2lcor", C.c_void_p),
                ("lir", C.c_void_p), ("fpah", C.c_void_p), ("mdust", C.c_void_p), ("igm_abs", C.c_void_p),
                ("vdisp", C.c_void_p), ("ir_sed", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_trapz.restype = C.c_double
        L.orc_trapz.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_trapz_range.restype = C.c_double
        L.orc_trapz_range.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.orc_filter_rlam.restype = C.c_double
        L.orc_filter_rlam.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_adjust_filter.restype = C.c_int64
        L.orc_adjust_filter.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_double)]
        L.orc_opt_lib_prepare.restype = C.c_int64
        L.orc_opt_lib_prepare.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_double, C.c_void_p, C.c_void_p]
        L.orc_merge_add.restype = C.c_int64
        L.orc_merge_add.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        L.orc_sed2flux.restype = C.c_double
        L.orc_sed2flux.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        for fn in (L.orc_get_flux, L.orc_get_flux_f64):
            fn.restype = None
            fn.argtypes = [C.POINTER(_Setup), C.POINTER(_Gal), C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_get_sed.restype = C.c_int64
        L.orc_get_sed.argtypes = [C.POINTER(_Setup), C.POINTER(_Gal), C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_totals.restype = None
        L.orc_totals.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
        L.orc_lsun2uJy_factor.restype = C.c_double
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


def trapz(x, y):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    return lib().orc_trapz(len(x), _p(x), _p(y))


def merge_add(x1, y1, x2, y2):
    x1, y1, x2, y2 = (np.ascontiguousarray(a, np.float64) for a in (x1, y1, x2, y2))
    x = np.empty(len(x1) + len(x2))
    y = np.empty(len(x1) + len(x2))
    n = lib().orc_merge_add(len(x1), _p(x1), _p(y1), len(x2), _p(x2), _p(y2), _p(x), _p(y))
    return x[:n].copy(), y[:n].copy()


def sed2flux(flam, fres, lam, sed, nodes="merged"):
    flam, fres, lam, sed = (np.ascontiguousarray(a, np.float64) for a in (flam, fres, lam, sed))
    return lib().orc_sed2flux(len(flam), _p(flam), _p(fres), len(lam), _p(lam), _p(sed), NODES[nodes])


def prepare_filters(raw, opts):
    """src/egg-gencat.cpp:316-378: optional adjustment, then sort by reference wavelength.
    raw: list of (lam, res) f64.  Returns (filters sorted, order, rlam f32 sorted)."""
    L = lib()
    fil, rl = [], []
    adjust = opts["trim_filters"] or opts["filter_flambda"] or opts["filter_photons"]
    for lam, res in raw:
        lam = np.array(lam, np.float64)
        res = np.array(res, np.float64)
        if adjust:
            r = C.c_double()
            n = L.orc_adjust_filter(len(lam), _p(lam), _p(res), int(opts["trim_filters"]),
                                    int(opts["filter_flambda"]), int(opts["filter_photons"]),
                                    int(opts["filter_normalize"]), C.byref(r))
            if n < 0:
                raise ValueError("filter has no usable data")
            lam, res, rlam = lam[:n].copy(), res[:n].copy(), r.value
        else:
            rlam = L.orc_filter_rlam(len(lam), _p(lam), _p(res))
        fil.append((lam, res))
        rl.append(rlam)
    lam32 = np.array(rl, np.float32)  # vec1f lambda (src/egg-gencat.cpp:368)
    order = np.argsort(lam32, kind="stable")
    return [fil[i] for i in order], order, lam32[order]


def prepare_opt_lib(opt_lib, opts):
    """src/egg-gencat.cpp:461-534.  Returns (lam [nsed][n] f32, sed [nsed][n] f32)."""
    lam = np.ascontiguousarray(opt_lib["lam"], np.float32)
    sed = np.ascontiguousarray(opt_lib["sed"], np.float32)
    use = np.ascontiguousarray(opt_lib["use"]).astype(np.uint8).ravel()
    nsed, nin = use.size, lam.shape[-1]
    imf = opts["opt_lib_imf"]
    if imf not in ("salpeter", "chabrier", "kroupa"):
        raise ValueError(f"unknown IMF '{imf}'")
    lo = np.empty(nsed * nin, np.float32)
    so = np.empty(nsed * nin, np.float32)
    n = lib().orc_opt_lib_prepare(nsed, nin, _p(lam), _p(sed), _p(use), int(imf != "salpeter"),
                                  float(opts["opt_lib_min_step"]), _p(lo), _p(so))
    if n < 0:
        raise ValueError("the provided optical SED library does not contain any valid SED")
    return lo[:nsed * n].reshape(nsed, n).copy(), so[:nsed * n].reshape(nsed, n).copy()


def _flatten(filters):
    off = np.zeros(len(filters) + 1, np.int64)
    for i, (l, _) in enumerate(filters):
        off[i + 1] = off[i] + len(l)
    lam = np.concatenate([f[0] for f in filters]) if filters else np.zeros(0)
    res = np.concatenate([f[1] for f in filters]) if filters else np.zeros(0)
    return off, np.ascontiguousarray(lam, np.float64), np.ascontiguousarray(res, np.float64)


class Oracle:
    """CPU restatement of the photometry stage for a fixed set of libraries, bands and options."""

    def __init__(self, opt_lib, ir_lib, bands_raw, rfbands_raw=(), line_lambda=None, **opts):
        o = dict(DEFAULT_OPTS)
        o.update(opts)
        self.opts = o
        self.filters, self.band_order, self.lambda_ = prepare_filters(list(bands_raw), o)
        self.rffilters, self.rfband_order, self.rflambda = prepare_filters(list(rfbands_raw), o)
        self.opt_lam, self.opt_sed = prepare_opt_lib(opt_lib, o)
        self.ir_kind = int(ir_lib["kind"])
        self.ir_lam = np.ascontiguousarray(ir_lib["lam"], np.float64)
        self.ir_sed = np.ascontiguousarray(ir_lib["sed"], np.float64)
        self.ir_pah = np.ascontiguousarray(ir_lib["pah"], np.float64) if ir_lib.get("pah") is not None else None
        if self.ir_lam.ndim == 1:  # 1-D generic library (src/egg-gencat.cpp:415-427)
            self.ir_lam, self.ir_sed = self.ir_lam[None, :], self.ir_sed[None, :]
        self.line_lambda = np.ascontiguousarray(line_lambda if line_lambda is not None else np.zeros(0), np.float32)
        self._foff, self._flam, self._fres = _flatten(self.filters)
        self._rfoff, self._rflam, self._rfres = _flatten(self.rffilters)
        s = _Setup()
        s.nopt_sed, s.nopt = self.opt_lam.shape
        s.opt_lam, s.opt_sed = _p(self.opt_lam), _p(self.opt_sed)
        s.ir_kind = self.ir_kind
        s.nir_sed, s.nir = self.ir_lam.shape
        s.ir_lam, s.ir_sed, s.ir_pah = _p(self.ir_lam), _p(self.ir_sed), _p(self.ir_pah)
        s.nband, s.foff, s.flam, s.fres = len(self.filters), _p(self._foff), _p(self._flam), _p(self._fres)
        s.nrf, s.rfoff, s.rflam, s.rfres = len(self.rffilters), _p(self._rfoff), _p(self._rflam), _p(self._rfres)
        s.nline, s.line_lambda = len(self.line_lambda), _p(self.line_lambda)
        s.apply_igm = int(o["igm"] not in ("none", "constant"))
        s.no_stellar, s.no_nebular = int(o["no_stellar"]), int(o["no_nebular"] or len(self.line_lambda) == 0)
        s.nodes_mode, s.gauss_mode = NODES[o["sed2flux_nodes"]], GAUSS[o["line_profile"]]
        s.emulate_f32 = int(o["emulate_f32"])
        self._s = s

    def _gal(self, g):
        keep = {}
        G = _Gal()
        n = len(g["z"])
        G.ngal = n
        for k in ("z", "d", "m2lcor", "lir", "fpah", "mdust", "igm_abs", "vdisp"):
            if k in g and g[k] is not None:
                keep[k] = np.ascontiguousarray(g[k], np.float32)
            else:
                keep[k] = np.zeros(n * (3 if k == "igm_abs" else 1), np.float32)
            setattr(G, k, _p(keep[k]))
        keep["ir_sed"] = np.ascontiguousarray(g.get("ir_sed", np.zeros(n)), np.uint64)
        G.ir_sed = _p(keep["ir_sed"])
        return G, keep

    def compute(self, g, nthreads: int = 0, f64: bool = False):
        """Both get_flux passes + totals (src/egg-gencat.cpp:2206-2239).  Returns dict of columns;
        with f64=True the component fluxes / magnitudes are returned unrounded (double)."""
        L = lib()
        G, keep = self._gal(g)
        n, nb, nrf = G.ngal, self._s.nband, self._s.nrf
        dt = np.float64 if f64 else np.float32
        fn = L.orc_get_flux_f64 if f64 else L.orc_get_flux
        out = {}
        nl = self._s.nline
        for comp, no_ir in (("bulge", True), ("disk", bool(self.opts["no_dust"]))):
            flux = np.zeros((n, nb), dt)
            rfm = np.zeros((n, nrf), dt)
            if comp == "bulge" and self.opts["no_stellar"]:  # src/egg-gencat.cpp:2206
                out["flux_bulge"], out["rfmag_bulge"] = flux, rfm
                continue
            m = np.ascontiguousarray(g[f"m_{comp}"], np.float32)
            osed = np.ascontiguousarray(g[f"opt_sed_{comp}"], np.uint64)
            ll = g.get(f"line_lum_{comp}")
            ll = np.ascontiguousarray(ll, np.float32) if (ll is not None and nl) else None
            fn(C.byref(self._s), C.byref(G), _p(m), _p(osed), _p(ll), _p(flux) if nb else None,
               _p(rfm) if nrf else None, int(no_ir), nthreads)
            out[f"flux_{comp}"], out[f"rfmag_{comp}"] = flux, rfm
        if f64:
            out["flux"] = out["flux_disk"] + out["flux_bulge"]
            with np.errstate(divide="ignore", invalid="ignore"):
                out["rfmag"] = -2.5 * np.log10(10 ** (0.4 * (23.9 - out["rfmag_disk"])) +
                                               10 ** (0.4 * (23.9 - out["rfmag_bulge"]))) + 23.9
        else:
            f = np.zeros((n, nb), np.float32)
            mt = np.zeros((n, nrf), np.float32)
            L.orc_totals(n * nb, _p(out["flux_disk"]), _p(out["flux_bulge"]), _p(f), n * nrf,
                         _p(out["rfmag_disk"]), _p(out["rfmag_bulge"]), _p(mt))
            out["flux"], out["rfmag"] = f, mt
        return out

    def get_sed(self, g, i, comp):
        """Observed-frame (lam, sed) f32 of one component as save_sed writes it."""
        G, keep = self._gal(g)
        m = np.ascontiguousarray(g[f"m_{comp}"], np.float32)
        osed = np.ascontiguousarray(g[f"opt_sed_{comp}"], np.uint64)
        ll = g.get(f"line_lum_{comp}")
        ll = np.ascontiguousarray(ll, np.float32) if (ll is not None and self._s.nline) else None
        cap = self.opt_lam.shape[1] + self.ir_lam.shape[1] + 8
        lam = np.zeros(cap, np.float32)
        sed = np.zeros(cap, np.float32)
        no_ir = True if comp == "bulge" else bool(self.opts["no_dust"])
        n = lib().orc_get_sed(C.byref(self._s), C.byref(G), i, _p(m), _p(osed), _p(ll), int(no_ir), _p(lam), _p(sed))
        return lam[:n].copy(), sed[:n].copy()
