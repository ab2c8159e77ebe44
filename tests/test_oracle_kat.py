"""Oracle pinned against the reference's only known-answer data (KAT-1, doc/doc-egg-gencat.tex:32-55)
and against closed forms."""
import numpy as np
import pytest

from egg_b200 import synth
from oracle import oracle as orc

# list_bands output printed in the reference manual: name -> (ref-lam, FWHM), both as printed (vec1f -> string)
KAT1 = {
    "jwst-f070w": (0.695353, 0.168), "jwst-f090w": (0.902688, 0.208), "jwst-f115w": (1.15124, 0.271),
    "jwst-f150w": (1.50168, 0.337), "jwst-f200w": (1.99057, 0.471), "jwst-f277w": (2.78408, 0.729),
    "jwst-f356w": (3.55939, 0.831), "jwst-f444w": (4.44572, 1.1526),
    "2mass-Ks": (2.16848, 0.271), "flamingos-Ks": (2.15594, 0.308), "fourstar-Ks": (2.15584, 0.322),
    "hawki-Ks": (2.14845, 0.3241), "isaac-Ks": (2.16813, 0.266), "moircs-Ks": (2.15952, 0.273),
    "sofi-Ks": (2.16798, 0.259), "vista-Ks": (2.15276, 0.301), "wircam-Ks": (2.15923, 0.322),
}


@pytest.mark.parametrize("name", sorted(KAT1))
def test_kat1_reference_wavelength_and_fwhm(name):
    """src/egg-gencat.cpp:195-199: trlam = integrate(lam, lam*res) on the raw response; FWHM from res > max/2."""
    pack = synth.load_pack()
    lam = pack[f"filter/{name}/lam"].astype(np.float64)
    res = pack[f"filter/{name}/res"].astype(np.float64)
    rlam = np.float32(orc.trapz(lam, lam * res))
    idg = np.flatnonzero(res > 0.5 * res.max())
    fwhm = np.float32(lam[idg[-1]] - lam[idg[0]])
    ref_l, ref_w = KAT1[name]
    assert float(f"{rlam:.6g}") == pytest.approx(ref_l, abs=1.1e-6 * 10 ** np.floor(np.log10(ref_l)) * 10)
    assert f"{rlam:.6g}" == f"{ref_l:.6g}"
    assert f"{fwhm:.4g}" == f"{ref_w:.4g}"
    # the same number through the filter preparation used for the band sort
    assert orc.lib().orc_filter_rlam(len(lam), lam.ctypes.data, res.ctypes.data) == pytest.approx(float(rlam), rel=1e-6)


def test_flat_sed_through_unit_norm_filter_returns_the_flat_value():
    lam = np.linspace(0.1, 10, 400)
    fl = np.linspace(1.0, 2.0, 57)
    fr = np.exp(-0.5 * ((fl - 1.5) / 0.2) ** 2)
    fr /= orc.trapz(fl, fr)
    for nodes in ("merged", "filter"):  # every filter node is a quadrature node: the response integral is exact
        assert orc.sed2flux(fl, fr, lam, np.full_like(lam, 3.25), nodes) == pytest.approx(3.25, rel=1e-13)
    # on the SED nodes alone the response is under-sampled: close, not equal (why the node set matters)
    assert orc.sed2flux(fl, fr, lam, np.full_like(lam, 3.25), "sed") == pytest.approx(3.25, rel=5e-3)


def test_sed2flux_not_covered_is_nan():
    lam = np.linspace(1.0, 2.0, 11)
    fl = np.linspace(0.9, 1.5, 5)
    assert np.isnan(orc.sed2flux(fl, np.ones(5), lam, np.ones(11)))
    fl = np.linspace(1.5, 2.1, 5)
    assert np.isnan(orc.sed2flux(fl, np.ones(5), lam, np.ones(11)))


def test_linear_sed_times_linear_filter_matches_simpson_on_merged_nodes():
    """On the merged node set each cell holds a quadratic; trapezoid error there is exactly g''h^3/12."""
    lam = np.array([0.0, 1.0, 3.0])
    sed = np.array([1.0, 2.0, 0.0])
    fl = np.array([0.5, 2.0, 2.5])
    fr = np.array([0.0, 3.0, 1.0])
    # union nodes 0.5, 1, 2, 2.5 with R = 0,1,3,1 and S = 1.5,2,1,0.5
    x = np.array([0.5, 1.0, 2.0, 2.5]); g = np.array([0.0, 2.0, 3.0, 0.5])
    assert orc.sed2flux(fl, fr, lam, sed, "merged") == pytest.approx(np.trapezoid(g, x), rel=1e-14)
    # SED nodes + filter end points: 0.5, 1, 2.5
    x = np.array([0.5, 1.0, 2.5]); g = np.array([0.0, 2.0, 0.5])
    assert orc.sed2flux(fl, fr, lam, sed, "sed") == pytest.approx(np.trapezoid(g, x), rel=1e-14)
    # filter nodes only: 0.5, 2, 2.5
    x = np.array([0.5, 2.0, 2.5]); g = np.array([0.0, 3.0, 0.5])
    assert orc.sed2flux(fl, fr, lam, sed, "filter") == pytest.approx(np.trapezoid(g, x), rel=1e-14)


def test_merge_add_disjoint_identical_interleaved():
    """src/egg-utils.hpp:5-63"""
    x, y = orc.merge_add([1, 2, 3], [1, 1, 1], [4, 5], [2, 2])            # disjoint: values untouched
    assert list(x) == [1, 2, 3, 4, 5] and list(y) == [1, 1, 1, 2, 2]
    x, y = orc.merge_add([1, 2, 3], [1, 2, 3], [1, 2, 3], [10, 20, 30])    # identical: one node each, summed
    assert list(x) == [1, 2, 3] and list(y) == [11, 22, 33]
    x, y = orc.merge_add([1, 3], [10, 30], [2, 4], [5, 7])                 # interleaved
    assert list(x) == [1, 2, 3, 4]
    # 1: before curve 2 starts -> 10;  2: 5 + lerp(10,30 @2)=25;  3: 30 + lerp(5,7 @3)=36;  4: after curve 1 ends -> 7
    assert list(y) == [10, 25, 36, 7]


def test_igm_unit_factors_are_a_no_op_and_zero_below_600A():
    L = {"opt": synth.make_opt_lib(), "ir": synth.ce01_lib()}
    gal = synth.draw_galaxies(20, opt_lib=L["opt"], ir_lib=L["ir"], mmin=9.0, seed=3)
    gal["igm_abs"] = np.ones_like(gal["igm_abs"])
    gal["z"] = np.clip(gal["z"], 0.05, 1.2).astype(np.float32)  # keep every band redward of 600 A rest
    bands = synth.get_filters(synth.B4)
    a = orc.Oracle(L["opt"], L["ir"], bands, igm="inoue14", no_nebular=True).compute(gal, f64=True)
    b = orc.Oracle(L["opt"], L["ir"], bands, igm="none", no_nebular=True).compute(gal, f64=True)
    np.testing.assert_allclose(a["flux"], b["flux"], rtol=1e-13)


def test_lsun2ujy_constant_and_magnitude_zero_point():
    k = orc.lib().orc_lsun2uJy_factor()
    assert k == pytest.approx(1.0e32 * 3.839e26 / (2.9979e14 * 4 * np.pi * 3.0856e22 ** 2), rel=1e-15)
    assert k == pytest.approx(1.0703e-2, rel=1e-4)
    f = np.zeros(1, np.float32); m = np.zeros(1, np.float32)
    md = np.array([20.0], np.float32); mb = np.array([20.0], np.float32)
    orc.lib().orc_totals(0, None, None, None, 1, md.ctypes.data, mb.ctypes.data, m.ctypes.data)
    assert m[0] == pytest.approx(20.0 - 2.5 * np.log10(2.0), abs=1e-6)


def test_opt_lib_rebin_box_average_preserves_a_linear_sed():
    """src/egg-gencat.cpp:469-534 on a grid that is too fine between 0.3 and 0.5 um."""
    lam = np.concatenate([np.linspace(0.1, 0.3, 21)[:-1], np.arange(0.3, 0.5, 0.0002), np.linspace(0.5, 2.0, 31)])
    sed = (2.0 + 3.0 * lam).astype(np.float32)
    lib = {"lam": lam.astype(np.float32)[None, None, :], "sed": sed[None, None, :], "use": np.ones((1, 1), bool)}
    o = dict(orc.DEFAULT_OPTS)
    nl, ns = orc.prepare_opt_lib(lib, o)
    assert nl.shape[1] < lam.size
    np.testing.assert_allclose(ns[0], 2.0 + 3.0 * nl[0].astype(np.float64), rtol=3e-6)
    d = np.diff(nl[0].astype(np.float64))
    assert d.min() > 0
    o["opt_lib_imf"] = "chabrier"
    _, ns2 = orc.prepare_opt_lib(lib, o)
    np.testing.assert_allclose(ns2, ns * 10 ** -0.2, rtol=2e-7)
