"""ctypes mirror of include/eggphot.h (structs, enums, argument packing)."""
from __future__ import annotations

import ctypes as C

import numpy as np

FP32C, FP64 = 0, 1
IR_GENERIC, IR_CS17 = 0, 1
NODES = {"merged": 0, "sed": 1, "filter": 2}
LINE = {"integral": 0, "average": 1}
PRECISION = {"fp32c": FP32C, "fp32": FP32C, "fp64": FP64}
STATUS = {0: "OK", 1: "ERR_ARG", 2: "ERR_FILTER", 3: "ERR_LIBRARY", 4: "ERR_GRID", 5: "ERR_UNSUPPORTED",
          6: "ERR_INDEX", 7: "ERR_CUDA", 8: "ERR_IO"}


class Libs(C.Structure):
    _fields_ = [("nuv", C.c_uint32), ("nvj", C.c_uint32), ("nopt", C.c_uint32),
                ("opt_lam", C.c_void_p), ("opt_sed", C.c_void_p), ("opt_use", C.c_void_p),
                ("ir_kind", C.c_int32), ("nir_sed", C.c_uint32), ("nir", C.c_uint32),
                ("ir_lam", C.c_void_p), ("ir_sed", C.c_void_p), ("ir_pah", C.c_void_p)]


class Filters(C.Structure):
    _fields_ = [("nband", C.c_uint32), ("npts", C.c_void_p), ("lam", C.c_void_p), ("res", C.c_void_p)]


class Opts(C.Structure):
    _fields_ = [("precision", C.c_int32), ("igm", C.c_char_p),
                ("no_stellar", C.c_int32), ("no_dust", C.c_int32), ("no_nebular", C.c_int32),
                ("opt_lib_imf", C.c_char_p), ("opt_lib_min_step", C.c_double),
                ("filter_flambda", C.c_int32), ("filter_photons", C.c_int32), ("filter_normalize", C.c_int32),
                ("trim_filters", C.c_int32), ("sed2flux_nodes", C.c_int32), ("line_profile", C.c_int32),
                ("nline", C.c_uint32), ("line_lambda", C.c_void_p), ("device", C.c_int32),
                ("smem_table_budget", C.c_uint32)]


GAL_F32 = ("z", "d", "m_disk", "m_bulge", "m2lcor")
GAL_FIELDS = ("z", "d", "m_disk", "m_bulge", "m2lcor", "opt_sed_disk", "opt_sed_bulge", "ir_sed", "lir", "fpah",
              "mdust", "igm_abs", "vdisp", "line_lum_disk", "line_lum_bulge")
GAL_U64 = ("opt_sed_disk", "opt_sed_bulge", "ir_sed")


class Galaxies(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in GAL_FIELDS]


OUT_FIELDS = ("flux", "flux_disk", "flux_bulge", "rfmag", "rfmag_disk", "rfmag_bulge",
              "flux_disk_f64", "flux_bulge_f64", "rfmag_disk_f64", "rfmag_bulge_f64")


class Outputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OUT_FIELDS]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("ngroups", C.c_uint32), ("nodes_disk", C.c_uint32),
                ("nodes_bulge", C.c_uint32), ("nodes_ir", C.c_uint32), ("table_bytes", C.c_uint64),
                ("last_kernel_ms", C.c_double)]


DEFAULT_OPTS = dict(precision="fp32c", igm="inoue14", no_stellar=False, no_dust=False, no_nebular=False,
                    opt_lib_imf="salpeter", opt_lib_min_step=0.001,
                    filter_flambda=False, filter_photons=False, filter_normalize=True, trim_filters=False,
                    sed2flux_nodes="merged", line_profile="integral", device=0, smem_table_budget=0)


class Packed:
    """Keeps the numpy arrays behind a set of C structs alive."""

    def __init__(self):
        self.keep = []

    def arr(self, a, dtype):
        a = np.ascontiguousarray(a, dtype)
        self.keep.append(a)
        return a.ctypes.data

    def libs(self, opt_lib, ir_lib) -> Libs:
        L = Libs()
        if opt_lib is not None:
            lam = np.asarray(opt_lib["lam"])
            L.nuv, L.nvj, L.nopt = lam.shape
            L.opt_lam = self.arr(lam, np.float32)
            L.opt_sed = self.arr(opt_lib["sed"], np.float32)
            L.opt_use = self.arr(np.asarray(opt_lib["use"]).astype(np.uint8), np.uint8)
        if ir_lib is not None:
            lam = np.atleast_2d(np.asarray(ir_lib["lam"], np.float64))
            L.ir_kind = int(ir_lib.get("kind", 0))
            L.nir_sed, L.nir = lam.shape
            L.ir_lam = self.arr(lam, np.float64)
            L.ir_sed = self.arr(np.atleast_2d(ir_lib["sed"]), np.float64)
            if ir_lib.get("pah") is not None:
                L.ir_pah = self.arr(np.atleast_2d(ir_lib["pah"]), np.float64)
        return L

    def filters(self, raw) -> Filters:
        F = Filters()
        raw = list(raw)
        F.nband = len(raw)
        if not raw:
            return F
        npts = np.array([len(l) for l, _ in raw], np.uint32)
        lam_p = (C.c_void_p * len(raw))(*[self.arr(l, np.float64) for l, _ in raw])
        res_p = (C.c_void_p * len(raw))(*[self.arr(r, np.float64) for _, r in raw])
        self.keep += [npts, lam_p, res_p]
        F.npts = npts.ctypes.data
        F.lam = C.cast(lam_p, C.c_void_p)
        F.res = C.cast(res_p, C.c_void_p)
        return F

    def opts(self, line_lambda=None, **kw) -> Opts:
        o = dict(DEFAULT_OPTS)
        unknown = set(kw) - set(o)
        if unknown:
            raise TypeError(f"unknown option(s): {sorted(unknown)}")
        o.update(kw)
        O = Opts()
        O.precision = PRECISION[o["precision"]] if isinstance(o["precision"], str) else int(o["precision"])
        igm, imf = o["igm"].encode(), o["opt_lib_imf"].encode()
        self.keep += [igm, imf]
        O.igm, O.opt_lib_imf = igm, imf
        for k in ("no_stellar", "no_dust", "no_nebular", "filter_flambda", "filter_photons", "filter_normalize",
                  "trim_filters", "device", "smem_table_budget"):
            setattr(O, k, int(o[k]))
        O.opt_lib_min_step = float(o["opt_lib_min_step"])
        O.sed2flux_nodes = NODES[o["sed2flux_nodes"]]
        O.line_profile = LINE[o["line_profile"]]
        if line_lambda is not None and len(line_lambda):
            O.nline = len(line_lambda)
            O.line_lambda = self.arr(line_lambda, np.float32)
        return O

    def galaxies(self, g) -> "tuple[Galaxies, int]":
        """Host columns (numpy) -> struct of host pointers."""
        G = Galaxies()
        n = len(g["z"])
        for k in GAL_FIELDS:
            v = g.get(k)
            if v is None:
                continue
            setattr(G, k, self.arr(v, np.uint64 if k in GAL_U64 else np.float32))
        return G, n
