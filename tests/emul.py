"""ctypes driver for tests/host_emul/libeggphot_emul.so (TEST-ONLY host walk of the device algorithm)."""
import ctypes as C
import os
import subprocess

import numpy as np

from egg_b200 import _cabi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_emul", "emul.cpp")
_SO = os.path.join(_HERE, "host_emul", "libeggphot_emul.so")
_CSRC = os.path.join(_HERE, "..", "egg_b200", "csrc")


def build():
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("phot_tables.cpp", "phot_tables.hpp", "phot_core.cuh", "phot_lane.cuh")]
    if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-Wall", "-shared", "-o", _SO, _SRC,
                               os.path.join(_CSRC, "phot_tables.cpp")])
    return _SO


def compute(opt_lib, ir_lib, bands, rfbands, gal, line_lambda=None, **opts):
    L = C.CDLL(build())
    pk = A.Packed()
    libs, fb, fr = pk.libs(opt_lib, ir_lib), pk.filters(bands), pk.filters(rfbands)
    o = pk.opts(line_lambda=line_lambda, **opts)
    G, n = pk.galaxies(gal)
    nb, nrf = fb.nband, fr.nband
    fd, fbu = np.zeros((n, nb)), np.zeros((n, nb))
    rd, rb = np.zeros((n, nrf)), np.zeros((n, nrf))
    order = np.zeros(nb, np.uint32)
    lam = np.zeros(nb, np.float32)
    msg = C.create_string_buffer(256)
    L.emul_compute.restype = C.c_int
    rc = L.emul_compute(C.byref(libs), C.byref(fb), C.byref(fr), C.byref(o), C.byref(G), C.c_size_t(n),
                        C.c_void_p(fd.ctypes.data), C.c_void_p(fbu.ctypes.data), C.c_void_p(rd.ctypes.data),
                        C.c_void_p(rb.ctypes.data), C.c_void_p(order.ctypes.data), C.c_void_p(lam.ctypes.data), msg)
    if rc:
        raise RuntimeError(f"emul status {rc}: {msg.value.decode()}")
    return {"flux_disk": fd, "flux_bulge": fbu, "rfmag_disk": rd, "rfmag_bulge": rb, "band_order": order, "lambda": lam}
