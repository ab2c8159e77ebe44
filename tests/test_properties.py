"""Property recipes of egg-gencat (src/egg-gencat.cpp:1360-2022): include/eggprops.h.

CPU tier: the counter-based generator against the published Random123 known-answer vectors; the restatement
(oracle/props_oracle.py) with `no_random` against closed forms evaluated here by hand from the reference's lines;
the statistics of the random terms; independence of chunking; the C ABI's exported symbols; the loud failure without
a GPU.  GPU tier: the CUDA kernel against the restatement (same generator, same slots: identical realisations), and
the device-resident chain properties -> photometry against the host chain."""
import math
import os
import re

import numpy as np
import pytest

import cases
import egg_b200
from egg_b200 import properties, synth
from oracle import props_oracle as PO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sample(n=20000, seed=3):
    rng = np.random.default_rng(seed)
    z = rng.uniform(0.05, 9.0, n).astype(np.float32)
    m = rng.uniform(7.5, 11.9, n).astype(np.float32)
    return z, m, rng.random(n) < 0.25


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = PO.philox4x32_10(*c, *k)
        assert tuple(int(x) for x in got) == want


def test_no_random_reproduces_the_reference_formulas_by_hand():
    L = cases.libs()
    z = np.array([0.3, 1.2, 2.5, 4.0], dtype=np.float32)
    m = np.array([9.0, 10.5, 11.2, 10.0], dtype=np.float32)
    pas = np.array([False, True, False, False])
    o = PO.compute(z, m, pas, opt_lib=L["opt"], no_random=True)
    zz, mm = z.astype(float), m.astype(float)
    # bulge-to-total, gencat:1371-1372, and the component masses :1381-1387
    bt = np.where(pas, 10 ** (0.1 * (mm - 10) - 0.3), 10 ** (0.27 * (mm - 10) - 0.7))
    np.testing.assert_allclose(o["bt"], np.clip(bt, 0, 1), rtol=1e-14)
    np.testing.assert_allclose(o["m_bulge"], mm + np.log10(o["bt"]), rtol=1e-14)
    # main sequence :1461 and the passive branch :1466
    az = np.log10(1 + zz)
    ms = mm - 9.67 + 1.82 * az - 0.38 * np.maximum(0, mm - 9.59 - 2.22 * az) ** 2
    want = np.where(pas, np.minimum(ms, 0.5 * (mm - 11) + az - 0.6), ms)
    np.testing.assert_allclose(np.log10(o["sfr"]), want, rtol=1e-13)
    np.testing.assert_allclose(o["rsb"], want - ms, atol=1e-14)
    # IRX :1493-1494; SFR_IR + SFR_UV = SFR :1504-1505
    np.testing.assert_allclose(o["irx"], 10 ** ((0.45 * np.minimum(zz, 3) + 0.35) * (mm - 10.5) + 1.2), rtol=1e-13)
    np.testing.assert_allclose(o["sfrir"] + o["sfruv"], o["sfr"], rtol=1e-13)
    np.testing.assert_allclose(o["lir"], o["sfrir"] / 1.72e-10, rtol=1e-14)            # :1619
    # M/L correction :544-546: 0.15 up to z = 0.45, 0 from 1.3 to 6
    np.testing.assert_allclose(o["m2lcor"], [0.15, 0.15 * (1.3 - 1.2) / (1.3 - 0.45), 0.0, 0.0], atol=1e-7)
    # ce01 template index :1685-1690 (RSB = 0 for the actives): 26 below z = 1, then towards 40 at 1.5, 54 at 2.1 ...
    act = ~pas
    want_ir = np.round(np.interp(zz, [0.57, 1.0, 1.5, 2.1, 2.9, 4.0, 6.0], [26, 26, 40, 54, 53, 52, 52]))
    assert list(o["ir_sed"][act]) == list(want_ir[act].astype(int))
    # dust temperature :1596, :1606, :1608
    td = 32.9 + 4.6 * (zz - 2) - 0.77 + 10.1 * o["rsb"] - 1.5 * np.maximum(0, 2 - zz) * np.clip(mm - 10.7, 0, 1)
    np.testing.assert_allclose(o["tdust"], td, rtol=1e-13)
    # H-alpha from the SFR :1915 before reddening; Balmer decrement of case B :1919-1923 survives reddening only
    # through the Calzetti curve: check the unreddened ratio on a galaxy with A_V(lines) = 0
    k = PO.calzetti2000(PO.LINE_LAMBDA.astype(float))
    ha, hb = PO.LINES.index("halpha"), PO.LINES.index("hbeta")
    ratio = o["line_lum_disk"][:, hb] / o["line_lum_disk"][:, ha]
    np.testing.assert_allclose(ratio, 0.35 * 10 ** (-0.4 * o["avlines_disk"] * (k[hb] - k[ha])), rtol=1e-12)
    np.testing.assert_allclose(o["line_lum_disk"][:, ha], 10 ** (np.log10(o["sfr"]) + 7.519 - 0.4 * o["avlines_disk"] * k[ha]), rtol=1e-12)
    assert (o["line_lum_bulge"] == 0).all()                                          # :1886 and every line after it
    # the median IGM transmission falls with redshift and is a transmission
    ia = PO.compute(np.array([0.5, 3.0, 6.0]), np.full(3, 10.0), np.zeros(3, bool), opt_lib=L["opt"], no_random=True)["igm_abs"]
    assert (ia >= 0).all() and (ia <= 1).all() and (np.diff(ia[:, 2]) < 0).all()
    # luminosity distance of the std cosmology: 6.7 Gpc at z = 1 to better than a percent
    d1 = PO.compute(np.array([1.0]), np.array([10.0]), np.zeros(1, bool), opt_lib=L["opt"], no_random=True)["d"][0]
    assert abs(d1 / 6607.0 - 1) < 5e-3


def test_random_terms_have_the_reference_scatter():
    L = cases.libs()
    n = 200000
    z, m, pas = _sample(n)
    o = PO.compute(z, m, pas, opt_lib=L["opt"], seed=11)
    o0 = PO.compute(z, m, pas, opt_lib=L["opt"], no_random=True)
    act = ~pas
    # log-normal scatters: 0.2 dex on B/T (:1376, where the clamp does not bite), 0.4 dex on IRX (:1497)
    free = (o["bt"] < 1) & (o0["bt"] < 1)
    assert abs(np.std(np.log10(o["bt"][free] / o0["bt"][free])) - 0.2) < 0.01
    assert abs(np.std(np.log10(o["irx"] / o0["irx"])) - 0.4) < 0.005
    # main sequence: ms_disp = 0.3 for actives plus 3.3 % starbursts at +0.72 (:1473-1478), 0.4 for passives (:1474)
    dl = np.log10(o["sfr"] / o0["sfr"])
    burst = dl[act] > 0.72 + 3 * 0.3
    assert abs(np.mean(dl[act]) - 0.033 * 0.72) < 0.005 and burst.mean() < 0.033
    assert abs(np.std(dl[pas]) - 0.4) < 0.01
    # position angles uniform in (-90, 90) (:1411); axis ratios follow the tabulated densities (:1414-1436)
    assert abs(o["disk_angle"].mean()) < 0.5 and abs(o["disk_angle"].std() - 180 / math.sqrt(12)) < 0.5
    x = np.array(PO.BULGE_RATIO_X)
    for col, p in (("bulge_ratio", PO.BULGE_RATIO_P), ("disk_ratio", PO.DISK_RATIO_P)):
        p = np.array(p, float)
        fine = np.linspace(0, 1, 20001)
        pdf = np.interp(fine, x, p)
        mean = np.trapezoid(fine * pdf, fine) / np.trapezoid(pdf, fine)
        assert abs(o[col].mean() - mean) < 3e-3 and o[col].min() >= 0.1
    # half of the sub-dominant bulges of active galaxies are red (:1530-1533): red bulges sit at U-V ~ 1.85
    sub = act & (o["bt"] <= 0.6) & (m < 9.5)                 # low mass: the blue locus stays below U-V = 1.3
    assert abs((o["rfuv_bulge"][sub] - 0.4 * np.maximum((0.5 - z[sub]) / 0.5, 0) > 1.5).mean() - 0.5) < 0.02
    # the drawn IGM transmissions scatter around the median curve
    mid = np.abs(z - 3.0) < 0.2
    assert abs(np.median(o["igm_abs"][mid, 2]) - np.median(o0["igm_abs"][mid, 2])) < 0.02


def test_columns_do_not_depend_on_chunking():
    L = cases.libs()
    z, m, pas = _sample(5000)
    whole = PO.compute(z, m, pas, opt_lib=L["opt"], seed=5)
    a = PO.compute(z[:1234], m[:1234], pas[:1234], opt_lib=L["opt"], seed=5)
    b = PO.compute(z[1234:], m[1234:], pas[1234:], opt_lib=L["opt"], seed=5, first_index=1234)
    for k in whole:
        np.testing.assert_array_equal(whole[k], np.concatenate([a[k], b[k]]))
    other = PO.compute(z, m, pas, opt_lib=L["opt"], seed=6)
    assert (other["bt"] != whole["bt"]).mean() > 0.9


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "eggprops.h")).read()
    declared = set(re.findall(r"\b(eggprops_[a-z_]+)\s*\(", hdr))
    lib = egg_b200.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(properties.EXPORTS)
    names, lam = properties.line_table()
    assert names == PO.LINES and np.array_equal(lam, PO.LINE_LAMBDA)                  # gencat:1834-1838


def test_user_errors_and_no_gpu():
    L = cases.libs()
    with pytest.raises(egg_b200.EggphotError) as e:
        properties.Properties(L["opt"], igm="foo")
    assert e.value.code == 1
    with pytest.raises(egg_b200.EggphotError) as e:
        properties.Properties(L["opt"], ir_lib="dale14")
    assert e.value.code == 5 and "no calibration code" in str(e.value)               # gencat:1702
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(egg_b200.EggphotError) as e:
            properties.Properties(L["opt"])
        assert e.value.code == 7                                                      # no CPU fallback


OPTION_SETS = [dict(), dict(no_random=True), dict(igm="madau95", magdis_tdust=True, no_passive_lir=True, ms_disp=0.25),
               dict(ir_lib="cs17"), dict(ir_lib="m12", igm="none"), dict(no_nebular=True, seed=2**40 + 17)]


def _aux(L, o):
    if o.get("ir_lib") == "cs17":
        return {k: L["cs17"][k] for k in ("tdust", "lir_dust", "lir_pah", "l8_dust", "l8_pah")}
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("opts", OPTION_SETS)
def test_cuda_matches_the_restatement(opts):
    L = cases.libs()
    n = 200000
    z, m, pas = _sample(n, seed=9)
    first = 10**10 + 7                                                               # exercises the upper counter word
    ref = PO.compute(z, m, pas, opt_lib=L["opt"], ir_aux=_aux(L, opts), first_index=first, **opts)
    pr = properties.Properties(L["opt"], ir_aux=_aux(L, opts), **opts)
    out = pr.compute(z, m, pas, first_index=first)
    pr.close()
    assert set(out) == set(pr.columns())
    for k, v in out.items():
        want = ref[k]
        if v.dtype == np.uint64:
            np.testing.assert_array_equal(v, want, err_msg=k)
            continue
        w32 = np.asarray(want, np.float64).astype(np.float32)
        # both sides round a double to f32: they agree to an f32 ulp unless the doubles differ (libm vs CUDA, 1e-16)
        bad = ~((v == w32) | (np.abs(v.astype(np.float64) - w32) <= 2.4e-7 * np.abs(w32)) | (np.isnan(v) & np.isnan(w32)))
        assert bad.sum() == 0, (k, int(bad.sum()), v[bad][:3], w32[bad][:3])
        assert (v == w32).mean() > 0.999, k


@pytest.mark.gpu
def test_properties_feed_the_photometry_on_the_device():
    """eggprops_compute_device -> eggphot_compute_device with no host copy of the property columns in between."""
    import torch
    L = cases.libs()
    n = 5000
    z, m, pas = _sample(n, seed=4)
    bands = synth.get_filters(synth.B4 + ["spitzer-irac1", "herschel-pacs100"])
    pr = properties.Properties(L["opt"], seed=8)
    host = pr.compute(z, m, pas)
    ph = egg_b200.Photometry(L["opt"], L["ce01"], bands, line_lambda=PO.LINE_LAMBDA)
    gal = {k: host[k] for k in ("d", "m_disk", "m_bulge", "m2lcor", "opt_sed_disk", "opt_sed_bulge", "ir_sed", "lir", "igm_abs",
                                "vdisp", "line_lum_disk", "line_lum_bulge")}
    gal["z"] = z
    want = ph.compute(gal)
    # the same on the device
    dz, dm, dp = torch.from_numpy(z).cuda(), torch.from_numpy(m).cuda(), torch.from_numpy(pas.astype(np.uint8)).cuda()
    cols = {}
    for k in gal:
        if k == "z":
            continue
        shape = (n, properties.SHAPE[k]) if k in properties.SHAPE else (n,)
        cols[k] = torch.empty(shape, dtype=torch.int64 if k in properties.U64_COLS else torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    pr.compute_device(dz.data_ptr(), dm.data_ptr(), dp.data_ptr(), n, {k: v.data_ptr() for k, v in cols.items()}, stream=st.cuda_stream)
    outs = {k: torch.empty((n, len(bands)), dtype=torch.float32, device="cuda") for k in ("flux", "flux_disk", "flux_bulge")}
    ptrs = {k: v.data_ptr() for k, v in cols.items()}
    ptrs["z"] = dz.data_ptr()
    ph.compute_device(ptrs, n, {k: v.data_ptr() for k, v in outs.items()}, stream=st.cuda_stream)
    torch.cuda.synchronize()
    for k in outs:
        np.testing.assert_array_equal(outs[k].cpu().numpy(), want[k])
    assert (want["flux"] > 0).mean() > 0.9
    pr.close()
    ph.close()


def test_kernel_and_restatement_hold_the_same_calibration_tables():
    """CPU tier cross-check of the constants the GPU tier compares through results: IGM quantile fits, line list,
    axis-ratio densities (the kernel source is parsed, nothing is executed)."""
    src = open(os.path.join(ROOT, "egg_b200", "csrc", "props_kernels.cu")).read()

    def table(name):
        body = re.search(name + r"[^=]*=\s*\{(.*?)\};", src, re.S).group(1)
        return np.array([float(x.rstrip("f")) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?f?", body)], dtype=np.float32)
    for name, ref in (("c_igm_z1", PO.IGM_Z1), ("c_igm_dz1", PO.IGM_DZ1), ("c_igm_z2", PO.IGM_Z2), ("c_igm_dz2", PO.IGM_DZ2), ("c_igm_q", PO.IGM_Q)):
        np.testing.assert_array_equal(table(name).astype(np.float64), np.asarray(ref).ravel(), err_msg=name)
    np.testing.assert_array_equal(table("h_line_lambda"), PO.LINE_LAMBDA)
    np.testing.assert_array_equal(table("kBulgeP"), np.array(PO.BULGE_RATIO_P, np.float32))
    np.testing.assert_array_equal(table("kDiskP"), np.array(PO.DISK_RATIO_P, np.float32))
    np.testing.assert_array_equal(table("kRatioX"), np.array(PO.BULGE_RATIO_X, np.float32))
    assert re.findall(r'"([a-z0-9_]+)"', re.search(r"h_line_name[^=]*=\s*\{(.*?)\};", src, re.S).group(1)) == PO.LINES


def test_ir_library_calibrations_without_scatter():
    """The three IR-library branches of src/egg-gencat.cpp:1680-1725 on hand-picked galaxies."""
    L = cases.libs()
    z = np.array([0.0125, 0.8125, 2.0, 2.635, 5.0])
    m = np.full(5, 10.0)
    act = np.zeros(5, bool)
    # m12: the template index follows the library's redshift grid (:1692-1697); beyond it the last segment is extended
    o = PO.compute(z, m, act, opt_lib=L["opt"], ir_lib="m12", no_random=True)
    assert list(o["ir_sed"][:4]) == [0, 3, 6, 7] and o["ir_sed"][4] == 7
    assert (o["fpah"] == 0.04).all() and np.allclose(o["mdust"], o["lir"] / 1e3)       # placeholders, :1722-1724
    # cs17: index from the dust temperature (:1699), exact f_PAH and M_dust from the library integrals (:1708-1719)
    cs = L["cs17"]
    aux = {k: cs[k] for k in ("tdust", "lir_dust", "lir_pah", "l8_dust", "l8_pah")}
    o = PO.compute(z, m, act, opt_lib=L["opt"], ir_lib="cs17", ir_aux=aux, no_random=True)
    want = np.clip(np.round(np.interp(o["tdust"], cs["tdust"], np.arange(len(cs["tdust"])))), 0, len(cs["tdust"]) - 1)
    inside = (o["tdust"] >= cs["tdust"][0]) & (o["tdust"] <= cs["tdust"][-1])
    assert inside.any() and list(o["ir_sed"][inside]) == list(want[inside].astype(int))
    i = o["ir_sed"].astype(int)
    fp = np.clip(1 / (1 - (cs["lir_pah"][i] - cs["l8_pah"][i] * o["ir8"]) / (cs["lir_dust"][i] - cs["l8_dust"][i] * o["ir8"])), 0, 1)
    np.testing.assert_allclose(o["fpah"], fp, rtol=1e-13)
    np.testing.assert_allclose(o["mdust"] * (cs["lir_dust"][i] * (1 - fp) + cs["lir_pah"][i] * fp), o["lir"], rtol=1e-13)
    # a single-template library (:1682)
    assert (PO.compute(z, m, act, opt_lib=L["opt"], ir_lib="single", no_random=True)["ir_sed"] == 0).all()
