"""Host-side mirror of egg-gencat's photometry stage on top of the C ABI (libeggphot.so).

``Photometry`` plays the role of the block ``if (!no_flux) {...}`` of src/egg-gencat.cpp:2024-2240:
construct it from the libraries, band lists and options ``vif_main`` has prepared at that point
(option names = the reference's command-line keys, src/egg-gencat.cpp:137-149), call ``compute``
with the per-galaxy columns of the ``out`` struct (src/egg-gencat.cpp:551-603) and get the
``flux*`` / ``rfmag*`` column groups back, band axis in the reference's sorted order
(src/egg-gencat.cpp:367-378).

All computation happens in the CUDA library; there is no fallback: without libeggphot.so or
without a B200-class device construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi as A

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libeggphot.so")
_lib = None


class EggphotError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"eggphot {A.STATUS.get(code, code)}: {msg}")
        self.code = code


def load_library(path: str = None):
    """dlopen libeggphot.so and declare its entry points.  Raises if it is missing (build it with
    ``make -C egg_b200/csrc`` or ``__graft_entry__.build()``).  EGGPHOT_LIB names another build of the same library
    (tools/build_variant.sh: kernel tuning experiments)."""
    global _lib
    if _lib is not None:
        return _lib
    if path is None:
        path = os.environ.get("EGGPHOT_LIB") or _LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: the CUDA extension has not been built; there is no CPU fallback")
    L = C.CDLL(path)
    vp, sz = C.c_void_p, C.c_size_t
    L.eggphot_create.restype = C.c_int
    L.eggphot_create.argtypes = [C.POINTER(A.Libs), C.POINTER(A.Filters), C.POINTER(A.Filters), C.POINTER(A.Opts),
                                 C.POINTER(vp)]
    L.eggphot_destroy.restype = None
    L.eggphot_destroy.argtypes = [vp]
    L.eggphot_last_error.restype = C.c_char_p
    L.eggphot_last_error.argtypes = [vp]
    L.eggphot_nband.restype = C.c_uint32
    L.eggphot_nband.argtypes = [vp, C.c_int]
    L.eggphot_band_order.restype = C.c_int
    L.eggphot_band_order.argtypes = [vp, C.c_int, vp, vp]
    L.eggphot_compute.restype = C.c_int
    L.eggphot_compute.argtypes = [vp, C.POINTER(A.Galaxies), sz, C.POINTER(A.Outputs)]
    L.eggphot_compute_device.restype = C.c_int
    L.eggphot_compute_device.argtypes = [vp, C.POINTER(A.Galaxies), sz, C.POINTER(A.Outputs), vp]
    L.eggphot_synchronize.restype = C.c_int
    L.eggphot_synchronize.argtypes = [vp]
    L.eggphot_sed_size.restype = C.c_uint32
    L.eggphot_sed_size.argtypes = [vp, C.c_int]
    L.eggphot_get_seds.restype = C.c_int
    L.eggphot_get_seds.argtypes = [vp, C.POINTER(A.Galaxies), sz, sz, C.c_int, vp, vp]
    L.eggphot_get_stats.restype = C.c_int
    L.eggphot_get_stats.argtypes = [vp, C.POINTER(A.Stats)]
    _lib = L
    return L


EXPORTS = ("eggphot_create", "eggphot_destroy", "eggphot_last_error", "eggphot_nband", "eggphot_band_order",
           "eggphot_compute", "eggphot_compute_device", "eggphot_synchronize", "eggphot_sed_size",
           "eggphot_get_seds", "eggphot_get_stats", "eggphot_measure_fma_peak", "eggphot_get_grid", "eggphot_maglim",
           "eggphot_get_seds_packed_begin", "eggphot_get_seds_packed_wait", "eggphot_host_register", "eggphot_host_unregister")


class Photometry:
    """One context = one GPU, one set of libraries / bands / options."""

    def __init__(self, opt_lib, ir_lib, bands, rfbands=(), line_lambda=None, **opts):
        self._L = load_library()
        self._pk = A.Packed()
        libs = self._pk.libs(opt_lib, ir_lib)
        fb = self._pk.filters(bands)
        fr = self._pk.filters(rfbands)
        o = self._pk.opts(line_lambda=line_lambda, **opts)
        self._ctx = C.c_void_p()
        rc = self._L.eggphot_create(C.byref(libs), C.byref(fb), C.byref(fr), C.byref(o), C.byref(self._ctx))
        if rc:
            msg = self._L.eggphot_last_error(self._ctx).decode() if self._ctx else "allocation failed"
            if self._ctx:
                self._L.eggphot_destroy(self._ctx)
                self._ctx = None
            raise EggphotError(rc, msg)
        self.nband = self._L.eggphot_nband(self._ctx, 0)
        self.nrf = self._L.eggphot_nband(self._ctx, 1)
        self.band_order, self.lambda_ = self._order(0, self.nband)
        self.rfband_order, self.rflambda = self._order(1, self.nrf)
        self.nline = 0 if (line_lambda is None or opts.get("no_nebular")) else len(line_lambda)

    def _order(self, which, n):
        order = np.zeros(n, np.uint32)
        lam = np.zeros(n, np.float32)
        self._L.eggphot_band_order(self._ctx, which, order.ctypes.data, lam.ctypes.data)
        return order, lam

    def _check(self, rc):
        if rc:
            raise EggphotError(rc, self._L.eggphot_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.eggphot_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host columns in, host columns out (H2D + kernels + D2H inside) --------------------------------
    def compute(self, gal, f64: bool = False, out=None):
        """gal: dict of numpy columns (reference names).  Returns dict of flux/rfmag columns (f32).
        f64=True additionally returns the unrounded component columns (``*_f64``).
        ``out``: optional dict of preallocated (e.g. pinned) arrays to fill."""
        pk = A.Packed()
        G, n = pk.galaxies(gal)
        res = {} if out is None else out
        O = A.Outputs()
        shapes = {"flux": self.nband, "flux_disk": self.nband, "flux_bulge": self.nband,
                  "rfmag": self.nrf, "rfmag_disk": self.nrf, "rfmag_bulge": self.nrf}
        for k, nb in shapes.items():
            if nb == 0:
                res[k] = np.zeros((n, 0), np.float32)
                continue
            if k not in res:
                res[k] = np.empty((n, nb), np.float32)
            setattr(O, k, res[k].ctypes.data)
        if f64:
            for k in ("flux_disk", "flux_bulge", "rfmag_disk", "rfmag_bulge"):
                nb = self.nband if k.startswith("flux") else self.nrf
                if nb:
                    res[k + "_f64"] = np.empty((n, nb), np.float64)
                    setattr(O, k + "_f64", res[k + "_f64"].ctypes.data)
        self._check(self._L.eggphot_compute(self._ctx, C.byref(G), n, C.byref(O)))
        return res

    # -- device pointers in / out, asynchronous on a stream ------------------------------------------------
    def compute_device(self, gal_ptrs: dict, ngal: int, out_ptrs: dict, stream: int = 0):
        """gal_ptrs / out_ptrs: name -> device address (int).  Enqueues only."""
        G = A.Galaxies()
        for k, v in gal_ptrs.items():
            setattr(G, k, v)
        O = A.Outputs()
        for k, v in out_ptrs.items():
            setattr(O, k, v)
        self._check(self._L.eggphot_compute_device(self._ctx, C.byref(G), ngal, C.byref(O), C.c_void_p(stream)))

    def synchronize(self):
        self._check(self._L.eggphot_synchronize(self._ctx))

    def sed_size(self, component):
        """nodes of one component's SED ('bulge' | 'disk'): the length of each save_sed array."""
        return int(self._L.eggphot_sed_size(self._ctx, {"bulge": 0, "disk": 1}[component]))

    def get_seds(self, gal, first, count, component):
        """save_sed payload (src/egg-gencat.cpp:2161-2179): (lam, sed) f32 [count][n], component 'bulge'|'disk'."""
        comp = {"bulge": 0, "disk": 1}[component]
        n = self._L.eggphot_sed_size(self._ctx, comp)
        pk = A.Packed()
        G, _ = pk.galaxies(gal)
        lam = np.empty((count, n), np.float32)
        sed = np.empty((count, n), np.float32)
        self._check(self._L.eggphot_get_seds(self._ctx, C.byref(G), first, count, comp, lam.ctypes.data, sed.ctypes.data))
        return lam, sed

    def grid(self, comp):
        """Rest wavelengths of the prepared SED grid of a component ('bulge' | 'disk')."""
        ci = 0 if comp == "bulge" else 1
        self._L.eggphot_sed_size.restype = C.c_uint32
        x = np.zeros(int(self._L.eggphot_sed_size(self._ctx, ci)), np.float64)
        self._check(self._L.eggphot_get_grid(self._ctx, ci, C.c_void_p(x.ctypes.data)))
        return x

    def maglim(self, props, zb_lo, zb_hi, maglim, mmin_strict=4.0, mmax=12.0, ntrial=1000):
        """egg-gencat's mass-limit estimator (src/egg-gencat.cpp:903-969): the lowest mass worth generating per redshift
        slice for a magnitude limit in this context's single band.  The context must have been created with bands = [the
        selection band], no_dust, no_nebular, igm='none'; `props` is a Properties object created with no_random=True."""
        lo, hi = np.ascontiguousarray(zb_lo, np.float64), np.ascontiguousarray(zb_hi, np.float64)
        out = np.empty(len(lo), np.float32)
        self._L.eggphot_maglim.restype = C.c_int
        self._L.eggphot_maglim.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_double,
                                           C.c_double, C.c_uint32, C.c_void_p]
        self._check(self._L.eggphot_maglim(self._ctx, props._ctx, lo.ctypes.data, hi.ctypes.data, len(lo), float(maglim),
                                           float(mmin_strict), float(mmax), int(ntrial), out.ctypes.data))
        return out

    def fma_peak(self):
        """Measured FP32 / FP64 FMA peak of the context's GPU in TFLOP/s (the roofline denominators of bench.py)."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        self._L.eggphot_measure_fma_peak.restype = C.c_int
        self._check(self._L.eggphot_measure_fma_peak(self._ctx, C.byref(a), C.byref(b)))
        return {"fp32_tflops": a.value, "fp64_tflops": b.value}

    def stats(self):
        s = A.Stats()
        self._L.eggphot_get_stats(self._ctx, C.byref(s))
        return {k: getattr(s, k) for k, _ in A.Stats._fields_}
