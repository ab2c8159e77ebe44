"""Seeded SYNTHETIC inputs for the photometry path.

The reference's optical libraries (``opt_lib_fast*.fits``) and ``ir_lib_cs17.fits`` are
missing large blobs and its galaxy draws are stochastic, so benchmarks and parity tests
feed both the oracle and the CUDA path with:

* ``make_opt_lib``  - an optical library with the reference's column schema
  (``lam, sed [nuv][nvj][nlam] f32, use, av, bvj, buv``; src/egg-gencat.cpp:449-458);
* ``make_cs17_lib`` - an IR library with the cs17 schema (src/egg-gencat.cpp:401-409);
* ``draw_galaxies`` - the per-galaxy property columns the hot path reads
  (``z d m_disk m_bulge m2lcor opt_sed_* ir_sed lir fpah mdust igm_abs vdisp line_lum_*``),
  drawn from the reference's mass function and published recipes with fixed seeds.

Nothing here is required to match egg's RNG (BASELINE.json north_star): the same columns go
to both paths.  Everything produced here is labelled SYNTHETIC.
"""
from __future__ import annotations

import os

import numpy as np

DATA_PACK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "egg_share.npz")

# line list of src/egg-gencat.cpp:1834-1838 (names, rest wavelengths in um)
LINES = ["c2_157", "n2_205", "c1_609", "co10", "co21", "co32", "co43", "co54", "co65", "co76",
         "halpha", "hbeta", "hgamma", "hdelta", "n2_6583", "n2_6548", "o3_5007", "o3_4959", "o2_3727",
         "lyalpha"]
LINE_LAMBDA = np.array([157.71, 205.18, 609.14, 2600.9, 1300.4, 867.0, 650.27, 520.24, 433.57, 371.66,
                        0.65628, 0.48613, 0.43405, 0.41017, 0.65835, 0.65480, 0.50068, 0.49589, 0.37274,
                        0.12157], dtype=np.float32)

# band sets of SURVEY.md §8d
B4 = ["hst-f435w", "hst-f606w", "hst-f850lp", "hst-f160w"]
B23 = ["vimos-u", "hst-f435w", "hst-f606w", "hst-f775w", "hst-f814w", "hst-f850lp", "hst-f105w",
       "hst-f125w", "hst-f140w", "hst-f160w", "hawki-Ks", "spitzer-irac1", "spitzer-irac2",
       "spitzer-irac3", "spitzer-irac4", "spitzer-irs16", "spitzer-mips24", "herschel-pacs70",
       "herschel-pacs100", "herschel-pacs160", "herschel-spire250", "herschel-spire350",
       "herschel-spire500"]  # defaults, src/egg-gencat.cpp:125-130
B30 = B23 + ["cfht-u", "subaru-B", "subaru-V", "subaru-r", "subaru-i", "subaru-z", "vista-Y"]
B40 = (["cfht-u"] + [f"subaru-{b}" for b in "BgVriz"] +
       [f"subaru-IB{n}" for n in (427, 464, 484, 505, 527, 574, 624, 679, 709, 738, 767, 827)] +
       ["subaru-NB711", "subaru-NB816", "vista-Y", "vista-J", "vista-H", "vista-Ks",
        "hst-f814w", "hst-f125w", "hst-f160w", "spitzer-irac1", "spitzer-irac2", "spitzer-irac3",
        "spitzer-irac4", "spitzer-mips24", "herschel-pacs100", "herschel-pacs160",
        "herschel-spire250", "herschel-spire350", "herschel-spire500", "galex-nuv", "wircam-Ks"])
RF3 = ["johnson-U", "johnson-V", "2mass-J"]

_pack_cache = {}


def load_pack(path: str = DATA_PACK):
    """The packed reference data tables (filters, ce01, mass function); see tools/pack_share_data.py."""
    if path not in _pack_cache:
        _pack_cache[path] = np.load(path)
    return _pack_cache[path]


def get_filters(names, pack=None):
    """name list -> list of (lam f64, res f64) raw curves (promoting the four f32 curves)."""
    pack = pack or load_pack()
    out = []
    for n in names:
        out.append((np.ascontiguousarray(pack[f"filter/{n}/lam"], dtype=np.float64),
                    np.ascontiguousarray(pack[f"filter/{n}/res"], dtype=np.float64)))
    return out


def ce01_lib(pack=None):
    pack = pack or load_pack()
    return {"kind": 0, "lam": np.ascontiguousarray(pack["ce01/lam"], dtype=np.float64),
            "sed": np.ascontiguousarray(pack["ce01/sed"], dtype=np.float64), "pah": None}


def _planck_nu_lnu(lam_um, T):
    """lambda * B_lambda (arbitrary units), lam in um."""
    x = 14387.77 / (lam_um * T)
    return lam_um ** -4 / np.expm1(np.minimum(x, 700.0))


def opt_grid(hd: bool = False) -> np.ndarray:
    """SYNTHETIC BaSeL/BC03-like rest wavelength grid [um]: 1221 nodes; the
    ``hd`` variant samples 0.32-0.95 um at 1e-4 um so that opt_lib_min_step re-binning runs."""
    parts = [np.linspace(0.0091, 0.0905, 40),
             np.arange(0.0912, 0.32 - 1e-9, 0.001)]
    if hd:
        parts += [np.arange(0.32, 0.95 - 1e-9, 0.0001), np.arange(0.95, 1.0 - 1e-9, 0.002)]
    else:
        parts += [np.arange(0.32, 1.0 - 1e-9, 0.002)]
    parts += [np.arange(1.0, 2.5 - 1e-9, 0.005), np.arange(2.5, 10.0 - 1e-9, 0.04)]
    n = sum(len(p) for p in parts)
    target = 1221 if not hd else None
    ntail = (target - n) if target else 124
    parts.append(np.geomspace(10.0, 160.0, ntail))
    lam = np.concatenate(parts).astype(np.float32)
    assert np.all(np.diff(lam) > 0)
    return lam


def make_opt_lib(seed: int = 1, nuv: int = 30, nvj: int = 30, hd: bool = False):
    """SYNTHETIC optical library with the reference schema (src/egg-gencat.cpp:449-458).

    sed = smooth positive nu.L_nu per unit stellar mass [Lsun/Msun]: two-temperature Planck mix
    whose weights follow the (U-V, V-J) bin, a 4000 A break, a Lyman-limit drop and seeded
    low-amplitude wiggles.  ``use`` is true on an elliptical region (~60 % of the bins)."""
    rng = np.random.default_rng(seed)
    lam = opt_grid(hd)
    nl = len(lam)
    l64 = lam.astype(np.float64)
    buv = np.stack([np.linspace(0.0, 3.0, nuv + 1)[:-1], np.linspace(0.0, 3.0, nuv + 1)[1:]]).astype(np.float32)
    bvj = np.stack([np.linspace(-0.5, 2.5, nvj + 1)[:-1], np.linspace(-0.5, 2.5, nvj + 1)[1:]]).astype(np.float32)
    uvc = 0.5 * (buv[0] + buv[1])
    vjc = 0.5 * (bvj[0] + bvj[1])
    UV, VJ = np.meshgrid(uvc, vjc, indexing="ij")
    # ellipse along the UVJ sequence
    a = (UV - 1.3) * 0.8 + (VJ - 0.9) * 0.6
    b = -(UV - 1.3) * 0.6 + (VJ - 0.9) * 0.8
    use = (a / 1.45) ** 2 + (b / 0.62) ** 2 < 1.0
    av = np.clip(1.5 * (VJ - 0.2), 0.0, 4.0).astype(np.float32)
    sed = np.zeros((nuv, nvj, nl), np.float32)
    lamc = np.zeros((nuv, nvj, nl), np.float32)
    hot = _planck_nu_lnu(l64, 14000.0)
    hot /= hot.max()
    cool = _planck_nu_lnu(l64, 3800.0)
    cool /= cool.max()
    for iu in range(nuv):
        for iv in range(nvj):
            if not use[iu, iv]:
                continue
            w_hot = np.exp(-1.6 * uvc[iu])
            w_cool = 0.35 + 0.5 * (vjc[iv] + 0.5)
            s = w_hot * hot + w_cool * cool
            brk = 1.0 - (0.15 + 0.12 * uvc[iu]) / (1.0 + np.exp((l64 - 0.4) / 0.004))
            lyl = 0.02 + 0.98 / (1.0 + np.exp(-(l64 - 0.0912) / 0.0008))
            red = 10 ** (-0.4 * av[iu, iv] * 0.6 * np.clip(0.55 / l64, 0.0, 6.0) ** 1.1 * (l64 < 3.0))
            wig = 1.0 + 0.03 * np.sin(l64 * rng.uniform(40, 90) + rng.uniform(0, 6.28)) * (l64 < 2.5)
            s = s * brk * lyl * red * wig * (0.6 + 0.8 * rng.random())
            sed[iu, iv] = (s * 0.9).astype(np.float32)
            lamc[iu, iv] = lam
    return {"lam": lamc, "sed": sed, "use": use, "av": av, "bvj": bvj, "buv": buv,
            "synthetic": True}


def make_cs17_lib(seed: int = 2, nt: int = 64, nlam: int = 800):
    """SYNTHETIC IR library with the cs17 schema (src/egg-gencat.cpp:401-409): per unit dust mass,
    ``dust`` = modified black body (beta 1.5), ``pah`` = fixed Lorentzian-bump template."""
    tdust = np.linspace(15.0, 100.0, nt)
    lam1 = np.geomspace(1.0, 3000.0, nlam)
    lam = np.tile(lam1, (nt, 1))
    dust = np.zeros((nt, nlam))
    pah = np.zeros((nt, nlam))
    bumps = [(3.3, 0.05, 0.3), (6.2, 0.19, 1.0), (7.7, 0.7, 2.2), (8.6, 0.34, 0.8), (11.3, 0.4, 1.1),
             (12.7, 0.5, 0.6), (17.0, 1.1, 0.4)]
    p = np.zeros(nlam)
    for c, w, amp in bumps:
        p += amp * (w / 2) ** 2 / ((lam1 - c) ** 2 + (w / 2) ** 2)
    p += 0.05 * (lam1 / 10.0) ** 0.5 * np.exp(-lam1 / 40.0)
    for t in range(nt):
        mbb = lam1 ** -1.5 * _planck_nu_lnu(lam1, tdust[t])
        dust[t] = 120.0 * mbb / np.trapezoid(mbb / lam1, lam1) * (tdust[t] / 30.0) ** 5.5
        pah[t] = 8000.0 * p / np.trapezoid(p / lam1, lam1)

    def lint(y, lo, hi):
        m = (lam1 >= lo) & (lam1 <= hi)
        return np.trapezoid(y[:, m] / lam1[m], lam1[m], axis=1)

    return {"kind": 1, "lam": lam, "sed": dust, "pah": pah, "tdust": tdust,
            "lir_dust": lint(dust, 8, 1000), "lir_pah": lint(pah, 8, 1000),
            "l8_dust": lint(dust, 7, 9), "l8_pah": lint(pah, 7, 9), "synthetic": True}


# ---------------------------------------------------------------------------------------------
# galaxies


def _lumdist(z):
    """Luminosity distance [Mpc], H0=70, Om=0.3, OL=0.7 (doc/doc-egg-gencat.tex:125)."""
    zg = np.linspace(0.0, 16.0, 16001)
    ez = 1.0 / np.sqrt(0.3 * (1 + zg) ** 3 + 0.7)
    dc = np.concatenate([[0.0], np.cumsum(0.5 * (ez[1:] + ez[:-1]) * np.diff(zg))]) * (299792.458 / 70.0)
    return (1.0 + z) * np.interp(z, zg, dc)


def _dvdz_per_deg2(z):
    """Comoving volume per unit redshift per deg^2 [Mpc^3]."""
    dc = _lumdist(z) / (1.0 + z)
    ez = 1.0 / np.sqrt(0.3 * (1 + z) ** 3 + 0.7)
    return (np.pi / 180.0) ** 2 * dc ** 2 * (299792.458 / 70.0) * ez


def expected_counts(area, mmin, zmin=0.05, zmax=10.5, mmax=12.0, pack=None):
    """Expected N[zbin, mbin] for active and passive galaxies from mass_func_candels."""
    pack = pack or load_pack()
    zb, mb = pack["mf/zb"].astype(float), pack["mf/mb"].astype(float)
    act, pas = pack["mf/active"], pack["mf/passive"]
    zlo, zhi = np.clip(zb[0], zmin, zmax), np.clip(zb[1], zmin, zmax)
    mlo, mhi = np.clip(mb[0], mmin, mmax), np.clip(mb[1], mmin, mmax)
    vol = np.zeros(len(zlo))
    for i in range(len(zlo)):
        if zhi[i] > zlo[i]:
            zz = np.linspace(zlo[i], zhi[i], 33)
            vol[i] = np.trapezoid(_dvdz_per_deg2(zz), zz) * area
    dm = np.maximum(mhi - mlo, 0.0)
    return act * vol[:, None] * dm[None, :], pas * vol[:, None] * dm[None, :], (zlo, zhi, mlo, mhi)


def draw_galaxies(ngal=None, *, area=None, mmin=None, seed=42, opt_lib, ir_lib, zmin=0.05, zmax=10.5,
                  nebular=True, pack=None):
    """SYNTHETIC galaxy property columns (reference names, reference dtypes).

    Either ``ngal`` (count fixed, z-M distribution of area=0.08 mmin given) or ``area``+``mmin``
    (count = rounded expectation of the mass function).  Recipes follow the published relations
    the reference uses (B/T: src/egg-gencat.cpp:1371-1387; main sequence :1463-1483; IRX
    :1491-1506; L_IR :1610; ce01 index :1685-1690; lines :1885-1977) with seeded scatter."""
    rng = np.random.default_rng(seed)
    pack = pack or load_pack()
    na, npas, (zlo, zhi, mlo, mhi) = expected_counts(area if area else 0.08, mmin if mmin is not None else 9.0,
                                                      zmin, zmax, pack=pack)
    tot = na.sum() + npas.sum()
    if ngal is None:
        ngal = int(round(tot))
    p = np.concatenate([na.ravel(), npas.ravel()]) / tot
    cell = rng.choice(p.size, size=ngal, p=p)
    passive = cell >= na.size
    cell = np.where(passive, cell - na.size, cell)
    iz, im = np.unravel_index(cell, na.shape)
    z = (zlo[iz] + (zhi[iz] - zlo[iz]) * rng.random(ngal)).astype(np.float32)
    m = (mlo[im] + (mhi[im] - mlo[im]) * rng.random(ngal)).astype(np.float32)
    order = np.argsort(~passive, kind="stable")[::-1]  # actives first, as the reference orders them
    z, m, passive = z[order], m[order], passive[order]
    g = {"z": z, "m": m, "passive": passive}
    g["d"] = _lumdist(z.astype(np.float64)).astype(np.float32)
    bt = np.where(passive, 10 ** (0.1 * (m - 10.0) - 0.3), 10 ** (0.27 * (m - 10.0) - 0.7))
    bt = np.clip(bt * 10 ** (0.2 * rng.standard_normal(ngal)), 0.0, 1.0).astype(np.float32)
    g["bt"] = bt
    with np.errstate(divide="ignore"):
        mb_ = (m + np.log10(bt)).astype(np.float32)
        md_ = (m + np.log10(1.0 - bt)).astype(np.float32)
    g["m_bulge"] = np.where(np.isfinite(mb_), mb_, 0).astype(np.float32)
    g["m_disk"] = np.where(np.isfinite(md_), md_, 0).astype(np.float32)
    g["m2lcor"] = np.interp(z, [0.0, 0.45, 1.3, 6.0, 8.0], [0.15, 0.15, 0.0, 0.0, -0.6]).astype(np.float32)
    az = np.log10(1.0 + z)
    sfrms = m - 9.67 + 1.82 * az - 0.38 * np.maximum(0.0, m - 9.59 - 2.22 * az) ** 2
    sfr = np.where(passive, np.minimum(sfrms, 0.5 * (m - 11) + az - 0.6) + 0.4 * rng.standard_normal(ngal),
                   sfrms + 0.3 * rng.standard_normal(ngal) + 0.72 * (rng.random(ngal) < 0.033))
    rsb = sfr - sfrms
    sfr = 10 ** sfr
    irx = 10 ** ((0.45 * np.minimum(z, 3.0) + 0.35) * (m - 10.5) + 1.2 + 0.4 * rng.standard_normal(ngal))
    sfrir = sfr / (1.0 + 1.0 / irx)
    sfruv = sfr / (1.0 + irx)
    g["lir"] = (sfrir / 1.72e-10).astype(np.float32)
    g["sfr"] = sfr.astype(np.float32)
    # optical SED bins: blue locus for active disks, red for passive; bulges red w.p. 1/2
    use = np.asarray(opt_lib["use"])
    nuv, nvj = use.shape
    ids = np.flatnonzero(use.ravel())
    iu, iv = np.unravel_index(ids, use.shape)
    w_red = np.exp(-0.5 * (((iu / nuv * 3.0) - 1.85) / 0.35) ** 2 - 0.5 * (((iv / nvj * 3.0 - 0.5) - 1.25) / 0.3) ** 2) + 1e-3
    w_blue = np.exp(-0.5 * (((iu / nuv * 3.0) - 0.9) / 0.5) ** 2 - 0.5 * (((iv / nvj * 3.0 - 0.5) - 0.7) / 0.5) ** 2) + 1e-3
    pick_red = rng.choice(ids, size=ngal, p=w_red / w_red.sum())
    pick_blue = rng.choice(ids, size=ngal, p=w_blue / w_blue.sum())
    pick_blue2 = rng.choice(ids, size=ngal, p=w_blue / w_blue.sum())
    g["opt_sed_disk"] = np.where(passive, pick_red, pick_blue).astype(np.uint64)
    red_bulge = (bt > 0.6) | passive | (rng.random(ngal) < 0.5)
    g["opt_sed_bulge"] = np.where(red_bulge, pick_red, pick_blue2).astype(np.uint64)
    nir = ir_lib["lam"].shape[0]
    if ir_lib["kind"] == 1:
        tdust = 32.9 + 4.6 * (z - 2.0) + 10.1 * rsb + 2.0 * rng.standard_normal(ngal)
        fir = np.interp(tdust, ir_lib["tdust"], np.arange(nir))
        ir_sed = np.clip(np.rint(fir), 0, nir - 1).astype(np.int64)
        fpah = np.clip(0.04 + 0.02 * rng.standard_normal(ngal), 0.0, 1.0)
        mdust = g["lir"] / (ir_lib["lir_dust"][ir_sed] * (1 - fpah) + ir_lib["lir_pah"][ir_sed] * fpah)
        g["fpah"] = fpah.astype(np.float32)
        g["mdust"] = mdust.astype(np.float32)
    else:
        fir = np.interp(z, [0.57, 1.0, 1.5, 2.1, 2.9, 4.0, 6.0], [26, 26, 40, 54, 53, 52, 52]) + \
            15 * np.clip(rsb / 0.3, -2.0, 2.0)
        ir_sed = np.clip(np.rint(fir), 0, nir - 1).astype(np.int64)
        g["fpah"] = np.full(ngal, 0.04, np.float32)
        g["mdust"] = (g["lir"] / 1e3).astype(np.float32)
    g["ir_sed"] = ir_sed.astype(np.uint64)
    # IGM transmissions in the three UV windows (Madau95-like mean + line-of-sight scatter)
    t2 = np.exp(-3.6e-3 * (1.11 * (1 + z)) ** 3.46)
    t1 = np.exp(-(1.7e-3 + 1.2e-3 + 9.3e-4) * (0.97 * (1 + z)) ** 3.46)
    t0 = np.clip(np.exp(-0.25 * (1 + z) ** 1.5) * rng.random(ngal), 0, 1)
    sc = np.clip(1.0 + 0.15 * rng.standard_normal((ngal, 3)), 0.3, 1.7)
    g["igm_abs"] = np.clip(np.stack([t0, t1, t2], 1) * sc, 0.0, 1.0).astype(np.float32)
    # emission lines
    g["vdisp"] = (10 ** (0.12 * (m - 10) + 1.78 + 0.3 * rng.standard_normal(ngal))).astype(np.float32)
    ll = np.zeros((ngal, len(LINES)), np.float64)
    lsfr = np.log10(sfr)
    mh2 = 0.3 * g["mdust"].astype(np.float64) * 10 ** (2.23 + 0.04 * rng.standard_normal(ngal))
    ll[:, 0] = 10 ** (1.017 * (lsfr - 0.2) + 7.13 + 0.27 * rng.standard_normal(ngal))
    ll[:, 1] = 10 ** (1.053 * (lsfr - 0.2) + 5.31 + 0.26 * rng.standard_normal(ngal))
    ll[:, 2] = 0.00246 * mh2
    ll[:, 3] = 4.93e-5 * mh2
    for k, s in enumerate([2.3, 3.7, 4.5, 7.5, 8.5, 8.3]):
        ll[:, 4 + k] = ll[:, 3] * s
    ha = 10 ** (lsfr + 7.519 + 0.15 * rng.standard_normal(ngal))
    att = 10 ** (-0.4 * np.clip(0.5 + 0.5 * rng.standard_normal(ngal), 0, 6)[:, None] *
                 np.array([0.82, 1.17, 1.3, 1.36, 0.82, 0.82, 1.12, 1.13, 1.5, 0.0])[None, :])
    ll[:, 10] = ha
    ll[:, 11:14] = ha[:, None] * np.array([0.35, 0.163, 0.0862])[None, :]
    n2 = ha * 10 ** (-0.6 + 0.2 * rng.standard_normal(ngal))
    ll[:, 14] = n2
    ll[:, 15] = n2 / 3.0
    o3 = ll[:, 11] * 10 ** (0.2 + 0.3 * rng.standard_normal(ngal))
    ll[:, 16] = o3
    ll[:, 17] = o3 / 3.0
    ll[:, 18] = ha * 10 ** (-0.1 + 0.15 * rng.standard_normal(ngal))
    ll[:, 19] = 8.7 * ha * np.clip(0.445 * 10 ** (-1.2 + 0.4 * rng.standard_normal(ngal)), 0, 1)
    ll[:, 10:] *= att
    if not nebular:
        ll[:] = 0.0
    g["line_lum_disk"] = ll.astype(np.float32)
    g["line_lum_bulge"] = np.zeros_like(g["line_lum_disk"])  # src/egg-gencat.cpp:1886-1976
    g["id"] = np.arange(ngal, dtype=np.uint64)
    g["synthetic"] = True
    return g
