"""GPU tier (-m gpu): the CUDA path, called through the C ABI (libeggphot.so), against the oracle on
the same seeded inputs, against the committed golden vectors, and through size-independent
properties at the full BASELINE config size."""
import os

import numpy as np
import pytest

import cases
import egg_b200
from egg_b200 import synth
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(c, prec, **kw):
    ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], precision=prec, **c["opts"], **kw)
    out = ph.compute(c["gal"], f64=True)
    st = ph.stats()
    ph.close()
    assert st["kernel_launches"] >= 1  # the CUDA kernels ran; there is no other path
    return out


@pytest.mark.parametrize("name", sorted(cases.CASES))
@pytest.mark.parametrize("prec", ["fp32c", "fp64"])
def test_cuda_matches_oracle(name, prec):
    c = cases.make(name, ngal=200)
    O = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"])
    ref = O.compute(c["gal"], f64=True)
    ref32 = O.compute(c["gal"])
    out = _run(c, prec)
    for k in ("flux_disk", "flux_bulge"):
        e = cases.rel_err(out[k + "_f64"], ref[k], cases.FLOOR[prec])
        assert e.max() < cases.TOL[prec], f"{name} {k}: {e.max():.2e}"
        # the f32 catalog columns: same numbers after rounding (1 ulp slack for values on a rounding edge)
        normal = np.abs(ref32[k]) > 1e-36  # (below that the f32 column itself is denormal and carries fewer digits)
        e32 = cases.rel_err(out[k], ref32[k])[normal]
        assert e32.size == 0 or (e32.max() < 1.1e-5 if prec == "fp32c" else e32.max() < 1.3e-7)
        assert np.abs(out[k][~normal].astype(np.float64) - ref32[k][~normal]).max(initial=0.0) < 1e-40
    for k in ("rfmag_disk", "rfmag_bulge"):
        if ref[k].size:
            assert np.abs(out[k + "_f64"] - ref[k]).max() < cases.MAGTOL[prec], f"{name} {k}"
    # totals are formed from the f32 component columns exactly as at src/egg-gencat.cpp:2238-2239
    np.testing.assert_array_equal(out["flux"], out["flux_disk"] + out["flux_bulge"])
    if ref["rfmag"].size:
        md, mb = out["rfmag_disk"].astype(np.float64), out["rfmag_bulge"].astype(np.float64)
        tot = (-2.5 * np.log10(10 ** (0.4 * (23.9 - md)) + 10 ** (0.4 * (23.9 - mb))) + 23.9).astype(np.float32)
        np.testing.assert_allclose(out["rfmag"], tot, rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", ["b4_defaults", "b30_rf3_defaults", "b40_hd_cs17", "b23_no_dust", "b30_hd_noigm"])
def test_cuda_matches_golden(name):
    c = cases.make(name, ngal=256, seed=2024)
    g = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    for prec in ("fp32c", "fp64"):
        out = _run(c, prec)
        for k in ("flux_disk", "flux_bulge"):
            assert cases.rel_err(out[k + "_f64"], g[k], cases.FLOOR[prec]).max() < cases.TOL[prec]
        for k in ("rfmag_disk", "rfmag_bulge"):
            if g[k].size:
                assert np.abs(out[k + "_f64"] - g[k]).max() < cases.MAGTOL[prec]


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 33])
def test_ragged_and_empty_batches(n):
    c = cases.make("b4_defaults", ngal=max(n, 1))
    gal = {k: (v[:n] if isinstance(v, np.ndarray) else v) for k, v in c["gal"].items()}
    ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], line_lambda=c["lines"])
    out = ph.compute(gal)
    assert out["flux"].shape == (n, 4)
    if n:
        ref = Oracle(c["opt"], c["ir"], c["bands"], line_lambda=c["lines"]).compute(gal, f64=True)
        ok = np.abs(ref["flux_disk"]) > 1e-36
        assert cases.rel_err(out["flux_disk"], ref["flux_disk"])[ok].max() < 1.1e-5
    ph.close()


def test_uncovered_band_is_zero_and_bad_index_is_an_error():
    c = cases.make("b30_rf3_defaults", ngal=64)
    ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"])
    out = ph.compute(c["gal"])
    lowz = c["gal"]["z"] < 5.0
    assert np.all(out["flux_bulge"][lowz][:, -3:] == 0.0)          # src/egg-gencat.cpp:2198
    gal = dict(c["gal"])
    gal["ir_sed"] = gal["ir_sed"].copy()
    gal["ir_sed"][3] = 10_000
    with pytest.raises(egg_b200.EggphotError) as e:
        ph.compute(gal)
    assert e.value.code == 6
    unusable = int(np.flatnonzero(~np.asarray(c["opt"]["use"]).ravel())[0])
    gal = dict(c["gal"])
    gal["opt_sed_disk"] = gal["opt_sed_disk"].copy()
    gal["opt_sed_disk"][0] = unusable
    with pytest.raises(egg_b200.EggphotError):
        ph.compute(gal)
    out2 = ph.compute(c["gal"])                                     # the context stays usable
    np.testing.assert_array_equal(out2["flux"], out["flux"])
    ph.close()


def test_rfmag_non_finite_is_99():
    c = cases.make("b30_rf3_defaults", ngal=8)
    gal = dict(c["gal"])
    gal["m_bulge"] = np.full(8, -60.0, np.float32)  # 10^-60 Msun underflows the f32 flux: mag -> +inf -> 99 (:2156)
    ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"])
    out = ph.compute(gal)
    ref = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"]).compute(gal)
    np.testing.assert_allclose(out["rfmag_bulge"], ref["rfmag_bulge"], rtol=0, atol=2e-5)
    ph.close()


def test_save_sed_payload_matches_oracle():
    """src/egg-gencat.cpp:2161-2179: observed-frame lam then sed, f32, per component."""
    c = cases.make("b30_rf3_defaults", ngal=40)
    O = Oracle(c["opt"], c["ir"], c["bands"], line_lambda=c["lines"])
    ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], line_lambda=c["lines"], precision="fp64")
    for comp in ("bulge", "disk"):
        lam, sed = ph.get_seds(c["gal"], 5, 20, comp)
        for k in (0, 7, 19):
            rl, rs = O.get_sed(c["gal"], 5 + k, comp)
            assert lam.shape[1] == len(rl)
            np.testing.assert_allclose(lam[k], rl, rtol=1.3e-7)
            np.testing.assert_allclose(sed[k], rs, rtol=3e-7, atol=1e-30)
    ph.close()


def test_config5_save_sed_fp64_hd_noigm():
    """BASELINE.json configs[4]: opt_lib_fast_hd_noigm (hd library, igm=none), fp64 mode, save_sed: the spectra
    (src/egg-gencat.cpp:2164-2193) and the fluxes of the same context against the oracle."""
    c = cases.make("b30_hd_noigm", ngal=300)
    O = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"])
    ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], precision="fp64", **c["opts"])
    out = ph.compute(c["gal"], f64=True)
    ref = O.compute(c["gal"], f64=True)
    for k in ("flux_disk", "flux_bulge"):
        assert cases.rel_err(out[k + "_f64"], ref[k]).max() < cases.TOL["fp64"], k
    for comp in ("bulge", "disk"):
        lam, sed = ph.get_seds(c["gal"], 100, 64, comp)
        for k in (0, 31, 63):
            rl, rs = O.get_sed(c["gal"], 100 + k, comp)
            assert lam.shape[1] == len(rl)
            np.testing.assert_allclose(lam[k], rl, rtol=1.3e-7)
            np.testing.assert_allclose(sed[k], rs, rtol=3e-7, atol=1e-30)
    ph.close()


@pytest.mark.parametrize("nodes", ["sed", "filter"])
@pytest.mark.parametrize("name", ["b4_defaults", "b30_rf3_defaults", "b40_hd_cs17", "b23_no_stellar"])
def test_other_sed2flux_node_sets_match_oracle(name, nodes):
    """The two other readings of [vif] sed2flux's node set (`check sed2flux_nodes=` of the CLI): SED nodes
    plus the filter's end points, and filter nodes alone.  Same tolerances as the contract's merged set."""
    c = cases.make(name, ngal=300, seed=5)
    opts = dict(c["opts"], sed2flux_nodes=nodes)
    ref = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **opts).compute(c["gal"], f64=True)
    merged = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"]).compute(c["gal"], f64=True)
    assert cases.rel_err(merged["flux_disk"], ref["flux_disk"]).max() > 1e-6  # the node sets do differ
    for prec in ("fp32c", "fp64"):
        ph = egg_b200.Photometry(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], precision=prec, **opts)
        out = ph.compute(c["gal"], f64=True)
        ph.close()
        for k in ("flux_disk", "flux_bulge"):
            e = cases.rel_err(out[k + "_f64"], ref[k])
            assert e.max() < cases.TOL[prec], f"{name} {nodes} {prec} {k}: {e.max():.2e}"
        for k in ("rfmag_disk", "rfmag_bulge"):
            if ref[k].size:
                assert np.abs(out[k + "_f64"] - ref[k]).max() < cases.MAGTOL[prec], f"{name} {nodes} {prec} {k}"


def test_full_size_properties():
    """BASELINE configs[1] size (1.95e5 galaxies x 30 bands), no oracle: determinism, chunk invariance
    (host path in 2^18 chunks vs small catalogs), linearity in mass, additivity of components."""
    n = 195_000
    opt, ir = synth.make_opt_lib(), synth.ce01_lib()
    gal = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5, seed=42)
    bands = synth.get_filters(synth.B30)
    ph = egg_b200.Photometry(opt, ir, bands, line_lambda=synth.LINE_LAMBDA)
    a = ph.compute(gal)
    b = ph.compute(gal)
    for k in ("flux", "flux_disk", "flux_bulge"):
        np.testing.assert_array_equal(a[k], b[k])                   # bit-reproducible
        assert np.isfinite(a[k]).all() and (a[k] >= 0).all()
    np.testing.assert_array_equal(a["flux"], a["flux_disk"] + a["flux_bulge"])
    # any sub-catalog gives the same rows (no dependence on batch or chunk position)
    for first, count in ((0, 1000), (100_003, 777), (n - 50, 50)):
        sub = {k: (v[first:first + count] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in gal.items()}
        s = ph.compute(sub)
        np.testing.assert_array_equal(s["flux_disk"], a["flux_disk"][first:first + count])
        np.testing.assert_array_equal(s["flux_bulge"], a["flux_bulge"][first:first + count])
    # +1 dex of bulge mass = x10 bulge flux exactly up to f32 rounding; disk untouched
    g2 = dict(gal)
    g2["m_bulge"] = (gal["m_bulge"] + np.float32(1.0)).astype(np.float32)
    c2 = ph.compute(g2)
    np.testing.assert_array_equal(c2["flux_disk"], a["flux_disk"])
    m = a["flux_bulge"] > 1e-30
    ratio = c2["flux_bulge"][m].astype(np.float64) / a["flux_bulge"][m]
    # m_bulge - m2lcor is formed in f32 (src/egg-gencat.cpp:2106): +1 can move it by one ulp of ~10 -> 2.2e-6 in 10^m
    assert np.abs(ratio / 10.0 - 1.0).max() < 5e-6
    ph.close()


@pytest.mark.parametrize("name", ["b30_rf3_defaults", "b4_chabrier_flags"])
def test_large_sample_has_no_outliers(name):
    """4000 galaxies per case, both modes: rare lookup events (a node on a filter node, on the split node, on a
    bucket edge) show up only in large samples (tools/parity_report.py is the same check over every case)."""
    c = cases.make(name, ngal=4000, seed=77)
    ref = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"]).compute(c["gal"], f64=True)
    for prec in ("fp32c", "fp64"):
        out = _run(c, prec)
        for k in ("flux_disk", "flux_bulge"):
            assert cases.rel_err(out[k + "_f64"], ref[k], cases.FLOOR[prec]).max() < cases.TOL[prec], (prec, k)


def test_large_catalog_spot_check_against_the_oracle(monkeypatch):
    """2e6 galaxies x 30 bands through the host-memory entry point (three chunks of 6.7e5 galaxies, each worked off in
    sub-chunks of 4 400 batches with the SED-row buffer capped at 2 GiB: the paths configs[3] takes per GPU), 1e5 of them
    drawn at random checked against the oracle."""
    monkeypatch.setenv("EGGPHOT_LANE_SCRATCH_MB", "2048")
    n, ns = 2_000_000, 100_000
    opt, ir = synth.make_opt_lib(), synth.ce01_lib()
    gal = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5, seed=99)
    bands = synth.get_filters(synth.B30)
    ph = egg_b200.Photometry(opt, ir, bands, line_lambda=synth.LINE_LAMBDA)
    out = ph.compute(gal, f64=True)
    ph.close()
    pick = np.sort(np.random.default_rng(1).choice(n, ns, replace=False))
    sub = {k: (v[pick] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in gal.items()}
    ref = Oracle(opt, ir, bands, line_lambda=synth.LINE_LAMBDA).compute(sub, f64=True)
    for k in ("flux_disk", "flux_bulge"):
        e = cases.rel_err(out[k + "_f64"][pick], ref[k])
        assert e.max() < cases.TOL["fp32c"], f"{k}: {e.max():.2e} at {np.unravel_index(np.argmax(e), e.shape)}"
    np.testing.assert_array_equal(out["flux"], out["flux_disk"] + out["flux_bulge"])
