"""Host mirror of include/eggprops.h: the per-galaxy property recipes of egg-gencat (src/egg-gencat.cpp:1360-2022)
on the GPU.  Nothing here computes a property on the CPU: without the CUDA library or a device every call raises."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi as A
from .photometry import EggphotError, load_library

IR_LIB = {"single": 0, "ce01": 1, "m12": 2, "cs17": 3}
IGM = {"none": 0, "constant": 0, "madau95": 1, "inoue14": 2}
NLINE = 20

F32_COLS = ("d", "bt", "m_bulge", "m_disk", "disk_angle", "bulge_angle", "disk_ratio", "bulge_ratio", "disk_radius",
            "bulge_radius", "sfr", "rsb", "irx", "sfrir", "sfruv", "rfuv_disk", "rfvj_disk", "rfuv_bulge", "rfvj_bulge",
            "m2lcor")
OUT_FIELDS = F32_COLS + ("opt_sed_disk", "opt_sed_bulge", "av_disk", "av_bulge", "tdust", "ir8", "lir", "fpah", "mdust",
                         "ir_sed", "igm_abs", "avlines_disk", "avlines_bulge", "vdisp", "line_lum_disk", "line_lum_bulge")
U64_COLS = ("opt_sed_disk", "opt_sed_bulge", "ir_sed")
SHAPE = {"igm_abs": 3, "line_lum_disk": NLINE, "line_lum_bulge": NLINE}
EXPORTS = ("eggprops_create", "eggprops_destroy", "eggprops_last_error", "eggprops_compute", "eggprops_compute_device",
           "eggprops_line_lambda", "eggprops_line_name")


class PLibs(C.Structure):
    _fields_ = [("nuv", C.c_uint32), ("nvj", C.c_uint32), ("buv", C.c_void_p), ("bvj", C.c_void_p), ("use", C.c_void_p),
                ("av", C.c_void_p), ("ir_lib", C.c_int32), ("nir_sed", C.c_uint32), ("cs17_tdust", C.c_void_p),
                ("cs17_lir_dust", C.c_void_p), ("cs17_lir_pah", C.c_void_p), ("cs17_l8_dust", C.c_void_p),
                ("cs17_l8_pah", C.c_void_p)]


class POpts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("no_random", C.c_int32), ("no_stellar", C.c_int32),
                ("no_nebular", C.c_int32), ("no_passive_lir", C.c_int32), ("magdis_tdust", C.c_int32),
                ("ms_disp", C.c_double), ("igm", C.c_int32), ("h0", C.c_double), ("omega_m", C.c_double),
                ("omega_l", C.c_double), ("device", C.c_int32)]


class POut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OUT_FIELDS]


def _declare(L):
    if getattr(L, "_eggprops_declared", False):
        return L
    vp, sz = C.c_void_p, C.c_size_t
    L.eggprops_create.restype = C.c_int
    L.eggprops_create.argtypes = [C.POINTER(PLibs), C.POINTER(POpts), C.POINTER(vp)]
    L.eggprops_destroy.restype = None
    L.eggprops_destroy.argtypes = [vp]
    L.eggprops_last_error.restype = C.c_char_p
    L.eggprops_last_error.argtypes = [vp]
    L.eggprops_compute.restype = C.c_int
    L.eggprops_compute.argtypes = [vp, vp, vp, vp, sz, C.c_uint64, C.POINTER(POut)]
    L.eggprops_compute_device.restype = C.c_int
    L.eggprops_compute_device.argtypes = [vp, vp, vp, vp, sz, C.c_uint64, C.POINTER(POut), vp]
    L.eggprops_line_lambda.restype = C.POINTER(C.c_float)
    L.eggprops_line_lambda.argtypes = []
    L.eggprops_line_name.restype = C.c_char_p
    L.eggprops_line_name.argtypes = [C.c_uint32]
    L._eggprops_declared = True
    return L


def line_table():
    """(names, rest wavelengths [um]) of the 20 emission lines in column order (src/egg-gencat.cpp:1834-1838)."""
    L = _declare(load_library())
    lam = L.eggprops_line_lambda()
    return [L.eggprops_line_name(i).decode() for i in range(NLINE)], np.array([lam[i] for i in range(NLINE)], dtype=np.float32)


class Properties:
    """One context = one GPU, one optical / IR library and one set of options (the reference's command-line keys)."""

    def __init__(self, opt_lib, ir_lib="ce01", ir_aux=None, *, seed=42, no_random=False, no_stellar=False,
                 no_nebular=False, no_passive_lir=False, magdis_tdust=False, ms_disp=0.3, igm="inoue14", device=0):
        self._L = _declare(load_library())
        self._ctx = C.c_void_p()
        if igm not in IGM:
            raise EggphotError(1, f"unknown IGM prescription '{igm}'")       # src/egg-gencat.cpp:286-289
        if ir_lib not in IR_LIB:
            raise EggphotError(5, f"no calibration code available for the IR library '{ir_lib}'")   # :1702
        keep = []

        def ptr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        pl = PLibs()
        if opt_lib is not None:
            buv, bvj = np.asarray(opt_lib["buv"]), np.asarray(opt_lib["bvj"])
            pl.nuv, pl.nvj = buv.shape[1], bvj.shape[1]
            pl.buv, pl.bvj = ptr(buv, np.float32), ptr(bvj, np.float32)
            pl.use, pl.av = ptr(np.asarray(opt_lib["use"]) != 0, np.uint8), ptr(np.asarray(opt_lib["av"]).ravel(), np.float32)
        aux = ir_aux or {}
        pl.ir_lib = IR_LIB[ir_lib]
        pl.nir_sed = int(len(aux["tdust"]) if ir_lib == "cs17" else aux.get("nir_sed", {"ce01": 105, "m12": 8, "single": 1}[ir_lib]))
        if ir_lib == "cs17":
            for k in ("tdust", "lir_dust", "lir_pah", "l8_dust", "l8_pah"):
                setattr(pl, "cs17_" + k, ptr(aux[k], np.float64))
        po = POpts(seed=int(seed), no_random=int(no_random), no_stellar=int(no_stellar), no_nebular=int(no_nebular),
                   no_passive_lir=int(no_passive_lir), magdis_tdust=int(magdis_tdust), ms_disp=float(ms_disp), igm=IGM[igm],
                   h0=70.0, omega_m=0.3, omega_l=0.7, device=int(device))
        self.no_nebular = bool(no_nebular)
        rc = self._L.eggprops_create(C.byref(pl), C.byref(po), C.byref(self._ctx))
        if rc != 0:
            msg = self._L.eggprops_last_error(self._ctx).decode() if self._ctx else "allocation failed"
            self.close()
            raise EggphotError(rc, msg)

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.eggprops_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EggphotError(rc, self._L.eggprops_last_error(self._ctx).decode())

    def columns(self):
        skip = ("avlines_disk", "avlines_bulge", "vdisp", "line_lum_disk", "line_lum_bulge") if self.no_nebular else ()
        return [k for k in OUT_FIELDS if k not in skip]

    def compute(self, z, m, passive, first_index=0):
        """Host arrays in, dict of host arrays out (reference column names and types)."""
        z, m = np.ascontiguousarray(z, np.float32), np.ascontiguousarray(m, np.float32)
        pas = np.ascontiguousarray(np.asarray(passive) != 0, np.uint8)
        n = z.size
        out, po = {}, POut()
        for k in self.columns():
            shape = (n, SHAPE[k]) if k in SHAPE else (n,)
            out[k] = np.zeros(shape, dtype=np.uint64 if k in U64_COLS else np.float32)
            setattr(po, k, out[k].ctypes.data)
        self._check(self._L.eggprops_compute(self._ctx, z.ctypes.data, m.ctypes.data, pas.ctypes.data, n, int(first_index), C.byref(po)))
        return out

    def compute_device(self, z_ptr, m_ptr, passive_ptr, ngal, out_ptrs: dict, first_index=0, stream=0):
        """Device pointers in and out (ints), asynchronous on `stream`: the columns stay on the GPU for the photometry."""
        po = POut()
        for k, v in out_ptrs.items():
            if k not in OUT_FIELDS:
                raise KeyError(k)
            setattr(po, k, int(v))
        self._check(self._L.eggprops_compute_device(self._ctx, int(z_ptr), int(m_ptr), int(passive_ptr), int(ngal), int(first_index),
                                                    C.byref(po), C.c_void_p(int(stream))))
