"""egg-gencat's `maglim` mass-limit estimator (src/egg-gencat.cpp:903-969) pushes 1000 trial
masses per redshift slice through get_opt_sed / lsun2uJy / sed2flux in the selection band.  That loop is the same
path with the options no_dust, no_nebular, igm=none and the trial masses as the "galaxies": the C ABI serves it as
it is.  The recipes that draw the trial SEDs (gen_blue_uvj, get_sed_uvj, get_m2l_cor) are upstream and out of
scope; here they are replaced by seeded draws over the usable templates."""
import numpy as np
import pytest

import cases
import emul
from egg_b200 import synth
from oracle.oracle import Oracle


def _trial_set(nz=12, ntrial=1000, seed=3):
    L = cases.libs()
    opt, ir = L["opt"], L["ce01"]
    rng = np.random.default_rng(seed)
    zs = np.linspace(0.1, 6.0, nz).astype(np.float32)
    n = nz * ntrial
    use = np.flatnonzero(np.asarray(opt["use"]).ravel())
    g = synth.draw_galaxies(n, opt_lib=opt, ir_lib=ir, mmin=7.5, seed=seed)
    g["z"] = np.repeat(zs, ntrial)
    g["d"] = synth._lumdist(g["z"].astype(np.float64)).astype(np.float32)
    g["m_disk"] = np.tile(np.linspace(4.0, 13.0, ntrial), nz).astype(np.float32)       # rgen(mmin_strict, mmax, 1000)
    g["m2lcor"] = np.repeat(rng.uniform(0.0, 0.3, nz), ntrial).astype(np.float32)        # get_m2l_cor(tz)
    g["opt_sed_disk"] = rng.choice(use, n).astype(np.uint64)                             # get_sed_uvj(...)
    opts = {"no_dust": True, "no_nebular": True, "igm": "none"}
    return opt, ir, g, synth.get_filters(["hst-f160w"]), opts, nz, ntrial


def _mass_limits(flux, g, nz, ntrial, maglim=27.0):
    flim = 10 ** (0.4 * (23.9 - maglim))                                                 # mag2uJy
    out = np.empty(nz)
    for k in range(nz):
        f = flux[k * ntrial:(k + 1) * ntrial, 0]
        idf = np.flatnonzero(f - flim > 0)
        out[k] = max(4.0, g["m_disk"][k * ntrial + idf[0]] - 0.2) if idf.size else 13.0  # :957-962
    return out


def test_trial_masses_through_the_emulated_device_path():
    opt, ir, g, bands, opts, nz, nt = _trial_set(nz=4, ntrial=250)
    ref = Oracle(opt, ir, bands, **opts).compute(g, f64=True)
    out = emul.compute(opt, ir, bands, [], g, precision="fp32c", **opts)
    assert cases.rel_err(out["flux_disk"], ref["flux_disk"], 1e-5).max() < 1e-5
    np.testing.assert_array_equal(_mass_limits(out["flux_disk"], g, nz, nt), _mass_limits(ref["flux_disk"], g, nz, nt))


@pytest.mark.gpu
def test_mass_limit_per_slice_matches_the_oracle():
    import egg_b200
    opt, ir, g, bands, opts, nz, nt = _trial_set()
    ref = Oracle(opt, ir, bands, **opts).compute(g)
    ph = egg_b200.Photometry(opt, ir, bands, **opts)
    out = ph.compute(g)
    ph.close()
    assert cases.rel_err(out["flux_disk"], ref["flux_disk"], 1e-5).max() < 2e-5
    a, b = _mass_limits(out["flux_disk"], g, nz, nt), _mass_limits(ref["flux_disk"], g, nz, nt)
    np.testing.assert_array_equal(a, b)
    assert np.all(np.diff(a[a < 13.0]) >= -1.0)  # the limit rises with redshift up to the scatter of the trial SEDs


@pytest.mark.gpu
def test_maglim_entry_point_follows_the_reference_recipe():
    """eggphot_maglim (include/eggphot.h) against the same steps spelled out with the property oracle and the flux
    oracle: trial masses rgen(mmin_strict, mmax, 1000), colours of an active galaxy without scatter, template of those
    colours, m2lcor of the slice, flux in the selection band, first crossing - 0.2 (src/egg-gencat.cpp:928-962)."""
    import egg_b200
    from egg_b200 import properties
    from oracle import props_oracle
    L = cases.libs()
    opt, ir = L["opt"], L["ce01"]
    bands = synth.get_filters(["hst-f160w"])
    opts = {"no_dust": True, "no_nebular": True, "igm": "none"}
    edges = np.array([0.1, 0.3, 0.6, 1.0, 1.6, 2.4, 3.5, 5.0, 7.0])
    lo, hi = edges[:-1], edges[1:]
    mmin_strict, mmax, nt, maglim = 6.0, 12.0, 1000, 26.0
    ph = egg_b200.Photometry(opt, ir, bands, **opts)
    pr = properties.Properties(opt, no_random=True)
    got = ph.maglim(pr, lo, hi, maglim, mmin_strict, mmax, nt)
    # the same by hand
    nz = len(lo)
    z = np.repeat((0.5 * (lo + hi)).astype(np.float32), nt)
    m = np.tile((mmin_strict + (mmax - mmin_strict) * np.arange(nt) / (nt - 1)).astype(np.float32), nz)
    cols = pr.compute(z, m, np.zeros(len(z), bool))
    pr.close()
    ph.close()
    g = {"z": z, "d": cols["d"], "m2lcor": cols["m2lcor"], "m_disk": m, "opt_sed_disk": cols["opt_sed_disk"],
         "m_bulge": np.full(len(z), -60.0, np.float32), "opt_sed_bulge": cols["opt_sed_disk"]}
    ref = Oracle(opt, ir, bands, **opts).compute(g, f64=True)["flux_disk"][:, 0]
    flim = 10 ** (0.4 * (23.9 - maglim))
    want = np.empty(nz, np.float32)
    for k in range(nz):
        idf = np.flatnonzero(ref[k * nt:(k + 1) * nt] - flim > 0)
        want[k] = max(mmin_strict, float(m[k * nt + idf[0]]) - 0.2) if idf.size else mmax
    np.testing.assert_array_equal(got, want)
    assert got.min() >= mmin_strict and got.max() <= mmax and np.any(np.diff(got) > 0)
