"""CPU restatement (numpy, float64) of egg-gencat's per-galaxy property recipes, src/egg-gencat.cpp:1360-2022.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke()): the product path is the CUDA kernel of
egg_b200/csrc/props_kernels.cu and never imports this file.

What the reference does there: from (z, M*, passive) it draws every other column of the catalog with
element-wise formulas plus random scatter — bulge-to-total ratio, sizes, SFR, IRX, UVJ colours and the
optical SED they select, dust temperature / IR8 and the IR SED they select, IGM transmission, emission
lines.  These columns are the inputs of the photometry path (include/eggphot.h, eggphot_galaxies).

Parity statement: PARITY UNPINNED against the reference binary.  The reference draws its random numbers
from one sequential generator ([vif] randomn / randomu / random_coin / random_pdf on a std::mt19937
seeded once, consumed in array order), so its realisations depend on the catalog order and cannot be
reproduced without vif.  This restatement (and the CUDA kernel, identically) uses a counter-based
generator instead: Philox4x32-10 keyed by the seed, counter = (galaxy index, draw slot): a galaxy's
columns depend only on (seed, index, z, M*, passive).  What is checked: Philox against the published
known-answer vectors of Random123; every formula with `no_random` against the closed forms quoted from
the reference lines; the distributions (means and scatters) of the random terms; CUDA == this file.
[vif] functions restated from their documented meaning: interpolate (linear, linear extrapolation of the
end segments), clamp, random_pdf (inverse of the trapezoid CDF of the tabulated density), in_bin_open
(first and last bins open-ended), astar_find (nearest usable cell of the UVJ grid), propsize / lumdist
(flat LCDM H0=70, Om=0.3: vif's "std" cosmology).
"""
import numpy as np

# ---- Philox4x32-10 (Salmon et al. 2011, Random123): the counter-based generator both sides use --------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Arrays (or scalars) of 32-bit words in uint64 containers -> four arrays of 32-bit words."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & _MASK, np.uint64(k1) & _MASK
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


# draw slots: one Philox block per (galaxy, slot); a uniform uses word 0, a normal words 0 and 1 (Box-Muller)
(S_BT, S_ANGLE, S_BULGE_RATIO, S_BULGE_RADIUS, S_DISK_RATIO, S_DISK_RADIUS, S_SFR, S_BURST, S_IRX,
 S_DISK_UVJ_A, S_DISK_UVJ_VJ, S_DISK_UVJ_UV, S_RED_BULGE, S_BULGE_UVJ_A, S_BULGE_UVJ_VJ, S_BULGE_UVJ_UV,
 S_TDUST, S_IR8, S_IGM0, S_IGM1, S_IGM2, S_GDR, S_MH2, S_AVL_DISK, S_AVL_BULGE, S_FESC, S_VDISP,
 S_C2, S_N2_205, S_HALPHA, S_O3MN2, S_O3PN2, S_O2) = range(33)


class Draws:
    def __init__(self, seed, index):
        self.k0, self.k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        idx = np.asarray(index, dtype=np.uint64)
        self.i0, self.i1 = idx & _MASK, idx >> np.uint64(32)

    def _block(self, slot):
        return philox4x32_10(self.i0, self.i1, np.uint64(slot), np.uint64(0), self.k0, self.k1)

    def uniform(self, slot):
        r = self._block(slot)
        return (r[0].astype(np.float64) + 0.5) * 2.0 ** -32

    def normal(self, slot):
        r = self._block(slot)
        u1 = (r[0].astype(np.float64) + 0.5) * 2.0 ** -32
        u2 = (r[1].astype(np.float64) + 0.5) * 2.0 ** -32
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


# ---- [vif] helpers -------------------------------------------------------------------------------------------------
def interpolate(y, x, v):
    """[vif] interpolate(y, x, v): linear, the end segments extended linearly outside [x0, xn]."""
    y, x, v = np.asarray(y, float), np.asarray(x, float), np.asarray(v, float)
    i = np.clip(np.searchsorted(x, v, side="right") - 1, 0, x.size - 2)
    return y[i] + (y[i + 1] - y[i]) * (v - x[i]) / (x[i + 1] - x[i])


def e10(x):
    return 10.0 ** np.asarray(x, float)


def random_pdf(u, x, p):
    """[vif] random_pdf: a value of x distributed as the tabulated density p(x), from a uniform deviate u:
    inverse of the normalised trapezoid CDF, linear between the tabulated points."""
    x, p = np.asarray(x, float), np.asarray(p, float)
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (p[1:] + p[:-1]) * np.diff(x))])
    cdf /= cdf[-1]
    i = np.clip(np.searchsorted(cdf, u, side="right") - 1, 0, x.size - 2)
    w = np.where(cdf[i + 1] > cdf[i], (u - cdf[i]) / np.where(cdf[i + 1] > cdf[i], cdf[i + 1] - cdf[i], 1.0), 0.0)
    return x[i] + w * (x[i + 1] - x[i])


_GL_X, _GL_W = np.polynomial.legendre.leggauss(32)


def comoving_mpc(z, h0=70.0, om=0.3, ol=0.7):
    """Line-of-sight comoving distance [Mpc], 32-point Gauss-Legendre on [0, z] (the kernel uses the same nodes)."""
    z = np.asarray(z, float)
    zz = 0.5 * z[:, None] * (_GL_X[None, :] + 1.0)
    ok = 1.0 - om - ol
    e = np.sqrt(om * (1 + zz) ** 3 + ok * (1 + zz) ** 2 + ol)
    return 2.99792458e5 / h0 * 0.5 * z * np.sum(_GL_W[None, :] / e, axis=1)


def sed_map(use):
    """For every (u, v) bin of the UVJ grid the flat index of the closest usable SED ([vif] astar_find, gencat:688):
    smallest squared index distance, ties to the smaller u, then the smaller v."""
    use = np.asarray(use).astype(bool)
    nu, nv = use.shape
    uu, vv = np.nonzero(use)
    out = np.zeros((nu, nv), dtype=np.int64)
    for u in range(nu):
        for v in range(nv):
            d = (uu - u) ** 2 + (vv - v) ** 2
            k = np.lexsort((vv, uu, d))[0]
            out[u, v] = uu[k] * nv + vv[k]
    return out


def bin_open(x, bins):
    """Index of the bin holding x ([vif] in_bin_open, gencat:680-683): bins[0] lower and bins[1] upper edges,
    contiguous and increasing; the first bin also takes everything below it, the last everything above."""
    lo = np.asarray(bins[0], float)
    return np.clip(np.searchsorted(lo, x, side="right") - 1, 0, lo.size - 1)


# ---- IGM tables (Inoue+14 quantile fits, gencat:1776-1808; data of the prescription) --------------------------------
IGM_Z1 = np.array([[0.00000, -0.119247, -0.836013, 0.188466, 5.58061, 1.80948, 2.23333, 2.58904, 3.05070, 3.44932, 3.64749, 4.33884, 4.80831],
                   [4.71243, 4.51667, 5.13483, 5.51325, 6.02321, 6.11072, 3.96055, 6.16846, 6.20577, 4.11253, 6.25897, 6.30988, 6.30475],
                   [4.29922, 4.43374, 5.81153, 6.07670, 6.19915, 4.01990, 6.22758, 6.23932, 4.18000, 6.29191, 6.30732, 6.34660, 4.29822]])
IGM_DZ1 = np.array([[0.00000, 0.103763, 1.53267, 0.831935, 0.101257, 1.30343, 1.23153, 1.18386, 1.18305, 1.21584, 1.22993, 1.25208, 1.15790],
                    [1.65837, 1.78514, 1.81298, 1.76652, 1.33753, 1.19458, 2.24961, 1.11023, 1.07492, 2.16499, 1.03259, 0.998597, 1.00750],
                    [1.83693, 1.92457, 1.44038, 1.07901, 0.980853, 2.13450, 1.00173, 0.996720, 2.25818, 0.979706, 0.979425, 0.962433, 2.04183]])
IGM_Z2 = np.array([[0.00000, 0.00000, 8.03488, 4.25279, 0.989755, 6.97760, 0.00000, 0.00000, 7.46371, 5.79003, 5.31787, 4.35838, 4.12159],
                   [3.13640, 4.21302, 4.01985, 3.91207, 3.85061, 3.91916, 6.14354, 3.99942, 4.05655, 6.24326, 4.14120, 4.22585, 4.28847],
                   [5.85101, 5.60883, 3.88240, 3.87634, 3.98672, -9.52333, 4.10508, 4.13700, 6.26376, 4.22116, 4.24243, 4.30912, 9.86437]])
IGM_DZ2 = np.array([[0.00000, 0.00000, 0.209441, 0.0120493, 1.27333, 0.0810017, 0.00000, 0.00000, 5.04822, 3.71614, 3.30993, 2.40054, 2.10543],
                    [5.31391, 5.23251, 3.21628, 2.77969, 2.37366, 2.28242, 1.14531, 2.22511, 2.19319, 1.04437, 2.15259, 2.11453, 2.11378],
                    [9.48358, 6.52196, 2.79285, 2.47621, 2.31278, -1.12091, 2.28194, 2.27401, 0.988508, 2.24063, 2.23221, 2.20610, 0.233647]])
IGM_Q = np.array([0.0, 0.001, 0.023, 0.05, 0.16, 0.35, 0.50, 0.65, 0.84, 0.95, 0.977, 0.999, 1.0])
# the reference keeps these tables in single precision (vec2f / vec1f)
IGM_Z1, IGM_DZ1, IGM_Z2, IGM_DZ2, IGM_Q = (a.astype(np.float32).astype(np.float64) for a in (IGM_Z1, IGM_DZ1, IGM_Z2, IGM_DZ2, IGM_Q))

LINES = ["c2_157", "n2_205", "c1_609", "co10", "co21", "co32", "co43", "co54", "co65", "co76",
         "halpha", "hbeta", "hgamma", "hdelta", "n2_6583", "n2_6548", "o3_5007", "o3_4959", "o2_3727", "lyalpha"]
LINE_LAMBDA = np.array([157.71, 205.18, 609.14, 2600.9, 1300.4, 867.0, 650.27, 520.24, 433.57, 371.66,
                        0.65628, 0.48613, 0.43405, 0.41017, 0.65835, 0.65480, 0.50068, 0.49589, 0.37274, 0.12157], dtype=np.float32)

BULGE_RATIO_X = [0.0, 0.1, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.75, 0.85, 0.95, 1.0]
BULGE_RATIO_P = [0.0, 0.0, 165.0, 428.0, 773.0, 914.0, 1069.0, 1191.0, 1154.0, 1067.0, 639.0, 450.0]
DISK_RATIO_P = [0.0, 0.0, 313.0, 900.0, 1143.0, 1127.0, 1059.0, 898.0, 775.0, 548.0, 318.0, 200.0]


def erf(x):
    from math import erf as _erf
    return np.vectorize(_erf, otypes=[float])(np.asarray(x, float))


def erfc(x):
    from math import erfc as _erfc
    return np.vectorize(_erfc, otypes=[float])(np.asarray(x, float))


def _igm_inoue14(z, draws, no_random):
    """gencat:1765-1823: transmission in the three bands below Ly-alpha, drawn from the fitted quantile curves."""
    out = np.zeros((z.size, 3))
    for l in range(3):
        tq = np.zeros((z.size, IGM_Q.size))
        for q in range(IGM_Q.size):
            if IGM_DZ1[l, q] == 0.0:
                continue
            t = 0.5 * erfc(-(IGM_Z1[l, q] - z) / IGM_DZ1[l, q])   # = 0.5 (erf(x) + 1) without the cancellation
            if IGM_DZ2[l, q] != 0.0:
                t = t * 0.5 * erfc(-(IGM_Z2[l, q] - z) / IGM_DZ2[l, q])
            tq[:, q] = t
        if no_random:
            out[:, l] = tq[:, IGM_Q.size // 2]
        else:
            u = draws.uniform(S_IGM0 + l)
            i = np.clip(np.searchsorted(IGM_Q, u, side="right") - 1, 0, IGM_Q.size - 2)
            r = np.arange(z.size)
            out[:, l] = tq[r, i] + (tq[r, i + 1] - tq[r, i]) * (u - IGM_Q[i]) / (IGM_Q[i + 1] - IGM_Q[i])
    return out


def _igm_madau95(z):
    """gencat:1737-1763 (FAST's implementation of Madau 1995): mean of exp(-tau) over 100 wavelengths per band."""
    out = np.zeros((z.size, 3))
    k = np.arange(100) / 99.0
    tl = (1050.0 + (1170.0 - 1050.0) * k)[None, :] * (1.0 + z[:, None])
    out[:, 2] = np.mean(np.exp(-3.6e-3 * (tl / 1216.0) ** 3.46), axis=1)
    tl = (920.0 + (1015.0 - 920.0) * k)[None, :] * (1.0 + z[:, None])
    out[:, 1] = np.mean(np.exp(-1.7e-3 * (tl / 1026.0) ** 3.46 - 1.2e-3 * (tl / 972.5) ** 3.46 - 9.3e-4 * (tl / 950.0) ** 3.46), axis=1)
    return out


def blue_uvj(m, z, d, slots, no_random):
    """gen_blue_uvj, gencat:619-645."""
    a0 = 0.58 * erf(m - 10) + 1.39
    as_ = -0.34 + 0.3 * np.maximum(m - 10.35, 0.0)
    a = np.minimum(a0 + as_ * np.minimum(z, 3.3), 2.0)
    if not no_random:
        amp = 0.2 + (0.25 + 0.12 * np.clip((z - 0.5) / 2.0, 0, 1)) * np.maximum(1.0 - 2.0 * np.abs(m - (10.3 + 0.4 * erf(z - 1.5))), 0.0)
        a = a + amp * d.normal(slots[0])
    theta = np.arctan(0.65)
    vj, uv = a * np.cos(theta), 0.45 + a * np.sin(theta)
    if not no_random:
        vj, uv = vj + 0.15 * d.normal(slots[1]), uv + 0.15 * d.normal(slots[2])
    off = np.maximum((0.5 - z) / 0.5, 0.0)
    return uv + 0.4 * off, vj + 0.2 * off


def red_uvj(m, z, d, slots, no_random):
    """gen_red_uvj, gencat:647-668."""
    ms = 0.1 * (m - 11.0)
    if not no_random:
        ms = ms + 0.1 * d.normal(slots[0])
    ms = np.clip(ms, -0.1, 0.2)
    vj, uv = 1.25 + ms, 1.85 + ms * 0.88
    if not no_random:
        vj, uv = vj + 0.1 * d.normal(slots[1]), uv + 0.1 * d.normal(slots[2])
    off = np.maximum((0.5 - z) / 0.5, 0.0)
    return uv + 0.4 * off, vj + 0.2 * off


def calzetti2000(l):
    """gencat:1986-2002 (k(lambda)/Rv + 1 of Calzetti et al. 2000, floored at 0)."""
    l = np.asarray(l, float)
    irv = 1.0 / 3.33
    k = np.where(l <= 0.63, (2.659 * irv) * (-2.156 + 1.509 / l - 0.198 * l ** -2.0 + 0.011 * l ** -3.0) + 1.0,
                 (2.659 * irv) * (-1.857 + 1.040 / l) + 1.0)
    return np.maximum(k, 0.0)


def poly3(x, c):
    return c[0] + x * c[1] + x * x * c[2] + x * x * x * c[3]


def compute(z, m, passive, *, opt_lib, ir_lib="ce01", ir_aux=None, seed=42, first_index=0, no_random=False,
            no_stellar=False, no_nebular=False, no_passive_lir=False, magdis_tdust=False, ms_disp=0.3, igm="inoue14"):
    """All property columns of gencat:1360-2022 for galaxies (z, m, passive).  opt_lib: dict with buv, bvj [2][n],
    use [n][n], av [n*n]; ir_lib: "ce01" | "m12" | "cs17" | "single"; ir_aux (cs17): tdust, lir_dust, lir_pah,
    l8_dust, l8_pah; nir_sed via ir_aux["nir_sed"] for ce01 / m12."""
    z, m = np.asarray(z, float), np.asarray(m, float)
    passive = np.asarray(passive).astype(bool)
    n = z.size
    d = Draws(seed, first_index + np.arange(n, dtype=np.uint64))
    o = {}
    rnd = not no_random

    # bulge-to-total ratio (Lang+14), gencat:1367-1387
    bt = np.where(passive, e10(0.1 * (m - 10.0) - 0.3), e10(0.27 * (m - 10.0) - 0.7))
    if rnd:
        bt = bt * e10(0.2 * d.normal(S_BT))
    bt = np.clip(bt, 0.0, 1.0)
    with np.errstate(divide="ignore"):
        mb, md = m + np.log10(bt), m + np.log10(1.0 - bt)
    o["bt"] = bt
    o["m_bulge"] = np.where(np.isfinite(mb), mb, 0.0)
    o["m_disk"] = np.where(np.isfinite(md), md, 0.0)

    # morphology, gencat:1398-1449
    az = np.log10(1.0 + z)
    size_disk = np.where(z < 1.7, e10(0.41 - 0.22 * az + 0.2 * (np.minimum(m, 10.6) - 9.35)), e10(0.62 - 0.7 * az + 0.2 * (np.minimum(m, 10.6) - 9.35)))
    size_bulge = np.where(z < 0.5, e10(0.78 - 0.6 * az + 0.56 * (m - 11.25)), e10(0.90 - 1.3 * az + 0.56 * (m - 11.25)))
    o["disk_angle"] = (d.uniform(S_ANGLE) - 0.5) * 180.0
    o["bulge_angle"] = o["disk_angle"]
    o["bulge_ratio"] = random_pdf(d.uniform(S_BULGE_RATIO), BULGE_RATIO_X, BULGE_RATIO_P)
    o["disk_ratio"] = random_pdf(d.uniform(S_DISK_RATIO), BULGE_RATIO_X, DISK_RATIO_P)
    if rnd:
        size_bulge = size_bulge * e10(0.2 * d.normal(S_BULGE_RADIUS))
        size_disk = size_disk * e10(0.17 * d.normal(S_DISK_RADIUS))
    dc = comoving_mpc(z)
    psize = dc / (1.0 + z) * 1e3 * (np.pi / 180.0 / 3600.0)  # proper kpc per arcsec
    o["d"] = dc * (1.0 + z)                                  # luminosity distance [Mpc] (the path's `d` column)
    o["disk_radius"], o["bulge_radius"] = size_disk / psize, size_bulge / psize

    # SFR, gencat:1459-1484
    sfrms = m - 9.67 + 1.82 * az - 0.38 * np.maximum(0.0, m - 9.59 - 2.22 * az) ** 2
    lsfr = np.where(passive, np.minimum(sfrms, 0.5 * (m - 11) + az - 0.6), sfrms)
    if rnd:
        lsfr = lsfr + np.where(passive, 0.4, ms_disp) * d.normal(S_SFR)
        lsfr = lsfr + np.where(~passive & (d.uniform(S_BURST) < 0.033), 0.72, 0.0)
    rsb = lsfr - sfrms
    sfr = e10(lsfr)

    # IRX, gencat:1488-1506
    irx = e10((0.45 * np.minimum(z, 3.0) + 0.35) * (m - 10.5) + 1.2)
    if rnd:
        irx = irx * e10(0.4 * d.normal(S_IRX))
    if no_passive_lir:
        irx = np.where(passive, 0.0, irx)
    with np.errstate(divide="ignore"):
        sfrir = sfr / (1.0 + 1.0 / irx)
    sfruv = sfr / (1.0 + irx)
    o.update(sfr=sfr, rsb=rsb, irx=irx, sfrir=sfrir, sfruv=sfruv)

    # UVJ colours, gencat:1515-1546
    ub, vb = blue_uvj(m, z, d, (S_DISK_UVJ_A, S_DISK_UVJ_VJ, S_DISK_UVJ_UV), no_random)
    ur, vr = red_uvj(m, z, d, (S_DISK_UVJ_A, S_DISK_UVJ_VJ, S_DISK_UVJ_UV), no_random)
    o["rfuv_disk"], o["rfvj_disk"] = np.where(passive, ur, ub), np.where(passive, vr, vb)
    red_bulge = (bt > 0.6) | passive
    if rnd:
        red_bulge = red_bulge | (d.uniform(S_RED_BULGE) < 0.5)
    ub, vb = blue_uvj(m, z, d, (S_BULGE_UVJ_A, S_BULGE_UVJ_VJ, S_BULGE_UVJ_UV), no_random)
    ur, vr = red_uvj(m, z, d, (S_BULGE_UVJ_A, S_BULGE_UVJ_VJ, S_BULGE_UVJ_UV), no_random)
    o["rfuv_bulge"], o["rfvj_bulge"] = np.where(red_bulge, ur, ub), np.where(red_bulge, vr, vb)
    o["m2lcor"] = interpolate([0.15, 0.15, 0.0, 0.0, -0.6], [0.0, 0.45, 1.3, 6.0, 8.0], z)   # gencat:544-546

    # optical SEDs, gencat:1554-1571, 670-705
    if not no_stellar:
        smap = sed_map(opt_lib["use"])
        av = np.asarray(opt_lib["av"], float).ravel()
        for c in ("bulge", "disk"):
            iu = bin_open(o["rfuv_" + c].astype(np.float32), np.asarray(opt_lib["buv"], np.float32))
            iv = bin_open(o["rfvj_" + c].astype(np.float32), np.asarray(opt_lib["bvj"], np.float32))
            o["opt_sed_" + c] = smap[iu, iv].astype(np.uint64)
            o["av_" + c] = av[smap[iu, iv]]
    else:
        for c in ("bulge", "disk"):
            o["opt_sed_" + c] = np.full(n, np.iinfo(np.uint64).max, dtype=np.uint64)
            o["av_" + c] = np.full(n, np.nan)

    # IR properties, gencat:1584-1668
    tdust_ms = 32.9 + 4.60 * (z - 2.0) - 0.77
    tdust = tdust_ms.copy()
    if magdis_tdust:
        tdust = np.where(z > 1, 2.0 * (np.clip(z, 0.0, 2.0) - 1.0) + 27.0, tdust)
    tdust = tdust + 10.1 * rsb
    tdust = tdust - 1.5 * np.maximum(0.0, 2.0 - z) * np.clip(m - 10.7, 0.0, 1.0)
    ir8 = (4.08 + 3.29 * np.clip(z - 1.0, 0.0, 1.0)) * 0.81
    ir8 = ir8 * e10(0.66 * np.maximum(0.0, rsb))
    ir8 = ir8 * e10(-1.0 * np.clip(m - 10.0, -1, 0.0))
    if rnd:
        tdust = tdust + 0.12 * tdust_ms * d.normal(S_TDUST)
        ir8 = ir8 * e10(0.18 * d.normal(S_IR8))
    lir = sfrir / 1.72e-10
    ir8 = np.clip(ir8, 0.48, 27.5)
    fpah = np.clip(1.0 / (1.0 - (331 - 691 * ir8) / (193 - 6.98 * ir8)), 0.0, 1.0)
    lir = np.where(np.isfinite(lir), lir, 0.0)

    # IR SED, gencat:1680-1725
    aux = ir_aux or {}
    if ir_lib == "single":
        nir, fir = 1, np.zeros(n)
    elif ir_lib == "ce01":
        nir = int(aux.get("nir_sed", 105))
        fir = interpolate([26, 26, 40, 54, 53, 52, 52], [0.57, 1.0, 1.5, 2.1, 2.9, 4.0, 6.0], z) + 15 * np.clip(rsb / ms_disp, -2.0, 2.0)
    elif ir_lib == "m12":
        nir = int(aux.get("nir_sed", 8))
        fir = interpolate(np.arange(nir), [0.0125, 0.1625, 0.4625, 0.8125, 1.15, 1.525, 2.0, 2.635][:nir], z) + np.clip(rsb / ms_disp, -2.0, 2.0)
    elif ir_lib == "cs17":
        nir = len(aux["tdust"])
        fir = interpolate(np.arange(nir), aux["tdust"], tdust)
    else:
        raise ValueError("no calibration code available for the IR library")   # gencat:1702
    # [vif] round = std::round (halves away from zero); negative values are clamped to 0 either way
    ir_sed = np.clip(np.floor(fir + 0.5), 0, nir - 1).astype(np.int64)
    if ir_lib == "cs17":
        ld, lp = np.asarray(aux["lir_dust"], float)[ir_sed], np.asarray(aux["lir_pah"], float)[ir_sed]
        l8d, l8p = np.asarray(aux["l8_dust"], float)[ir_sed], np.asarray(aux["l8_pah"], float)[ir_sed]
        fpah = np.clip(1.0 / (1.0 - (lp - l8p * ir8) / (ld - l8d * ir8)), 0.0, 1.0)
        mdust = lir / (ld * (1.0 - fpah) + lp * fpah)
    else:
        fpah = np.full(n, 0.04)
        mdust = lir / 1e3
    o.update(tdust=tdust, ir8=ir8, lir=lir, fpah=fpah, mdust=mdust, ir_sed=ir_sed.astype(np.uint64))

    # IGM, gencat:1731-1824
    if igm == "inoue14":
        o["igm_abs"] = _igm_inoue14(z, d, no_random)
    elif igm == "madau95":
        o["igm_abs"] = _igm_madau95(z)
    else:
        o["igm_abs"] = np.zeros((n, 3))   # the column is resized to zeros and left alone (gencat:1731, "constant" / "none")

    # emission lines, gencat:1829-2022
    if not no_nebular:
        nl = len(LINES)
        L = {k: i for i, k in enumerate(LINES)}
        ld_ = np.zeros((n, nl))
        with np.errstate(divide="ignore", invalid="ignore"):
            lsfr10 = np.log10(sfr)
            mu32 = (m - 0.2) - 0.32 * (lsfr10 - 0.2)
            metal = np.where(mu32 < 10.36, 8.9 + 0.47 * (mu32 - 10), 9.07) - 8.69
            metal = np.where(z > 1, metal - 0.2 * np.clip(z - 1, 0.0, 1.0), metal)
            metal = np.clip(metal, -0.6, 1.0)
            gdr = 2.23 - metal
            if rnd:
                gdr = gdr + 0.04 * d.normal(S_GDR)
            mh2 = mdust * e10(gdr) * 0.3
            if rnd:
                mh2 = mh2 * e10(0.2 * d.normal(S_MH2))
            avd = 0.5 * (np.log10(sfr / sfruv) * 0.95 + o["av_disk"])
            avd = avd * interpolate([1.7, 1.3, 1.0, 1.0], [0, 1, 2, 100], z)
            avb = np.asarray(o["av_bulge"], float).copy()
            if rnd:
                avd, avb = avd + 0.1 * d.normal(S_AVL_DISK), avb + 0.1 * d.normal(S_AVL_BULGE)
            avd, avb = np.clip(avd, 0.0, 6.0), np.clip(avb, 0.0, 6.0)
            fesc = 0.445 * e10(-0.4 * (o["av_disk"] / 4.05) * 17.8)
            fesc = fesc * e10(interpolate([-1.2, -1.2, -0.8, 0, 0], [0.5, 1.2, 2.2, 3.0, 4.0], z))
            if rnd:
                fesc = fesc * e10(0.4 * d.normal(S_FESC))
            fesc = np.clip(fesc, 0.0, 1.0)
            vdisp = e10(0.12 * (m - 10) + 1.78)
            if rnd:
                vdisp = vdisp * e10(0.3 * d.normal(S_VDISP))
            ld_[:, L["c2_157"]] = e10(1.017 * (lsfr10 - 0.2) + 7.13) * (e10(0.27 * d.normal(S_C2)) if rnd else 1.0)
            ld_[:, L["n2_205"]] = e10(1.053 * (lsfr10 - 0.2) + 5.31) * (e10(0.26 * d.normal(S_N2_205)) if rnd else 1.0)
            ld_[:, L["c1_609"]] = 0.00246 * mh2
            ld_[:, L["co10"]] = 4.93e-5 * mh2
            for k, s in enumerate([2.3, 3.7, 4.5, 7.5, 8.5, 8.3]):
                ld_[:, L["co10"] + 1 + k] = ld_[:, L["co10"]] * s
            lha = e10(lsfr10 + 7.519) * (e10(0.15 * d.normal(S_HALPHA)) if rnd else 1.0)
            ld_[:, L["halpha"]] = lha
            for k, r in enumerate([0.35, 0.163, 0.0862]):
                ld_[:, L["halpha"] + 1 + k] = lha * r
            lo3mn2 = poly3(metal, [0.59518604, -2.7283054, -0.23229248, 2.0151786])
            if rnd:
                lo3mn2 = lo3mn2 + poly3(metal, [0.32970659, 0.043645415, -1.2231136, -2.4836792]) * d.normal(S_O3MN2)
            lo3pn2 = poly3(lo3mn2, [-0.69337981, 0.66460936, -0.37040100, 0.047935498])
            if rnd:
                lo3pn2 = lo3pn2 + poly3(lo3mn2, [0.086819227, -0.078283561, 0.093276637, -0.025069864]) * d.normal(S_O3PN2)
            n2off = interpolate([0.0, 0.0, 0.3, 0.3], [0.0, 1.0, 2.3, 2.4], z) * interpolate([1.0, 1.0, 0.0, 0.0], [9, 10.0, 10.5, 11.0], m)
            ln2 = lha * e10(0.5 * (lo3pn2 - lo3mn2) + n2off)
            ld_[:, L["n2_6583"]], ld_[:, L["n2_6548"]] = 0.75 * ln2, 0.25 * ln2
            lhb = ld_[:, L["hbeta"]]
            lo3 = lhb * e10(0.5 * (lo3pn2 + lo3mn2))
            ld_[:, L["o3_5007"]], ld_[:, L["o3_4959"]] = 0.75 * lo3, 0.25 * lo3
            o3hb = np.log10(lo3 / lhb)
            lo2 = lhb * e10(poly3(o3hb, [0.51869283, 0.29843257, -0.49610294, -0.11728400]))
            if rnd:
                lo2 = lo2 * e10(poly3(o3hb, [0.061403650, -0.0081335335, 0.078420795, -0.032656826]) * d.normal(S_O2))
            ld_[:, L["o2_3727"]] = lo2
            ld_[:, L["lyalpha"]] = 8.7 * lha * fesc
            dust = calzetti2000(LINE_LAMBDA.astype(float))
            dust[L["lyalpha"]] = 0.0
            ll = LINE_LAMBDA.astype(float)
            igl = np.ones((n, nl))
            ia = o["igm_abs"]
            for l in range(nl):
                if ll[l] < 0.0600: igl[:, l] = 0.0
                elif ll[l] < 0.0912: igl[:, l] = ia[:, 0]
                elif ll[l] < 0.1026: igl[:, l] = ia[:, 1]
                elif ll[l] < 0.1216: igl[:, l] = ia[:, 2]
            o["line_lum_disk"] = ld_ * e10(-0.4 * avd[:, None] * dust[None, :]) * igl
            o["line_lum_bulge"] = np.zeros((n, nl))
        o.update(avlines_disk=avd, avlines_bulge=avb, vdisp=vdisp)
    return o
