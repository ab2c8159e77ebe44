"""egg-gencat-b200, the stand-alone drop-in for the photometry stage (egg_b200/csrc/gencat_main.cpp).

CPU tier: the command-line contract of egg-gencat that needs no device — `list_bands` against the
reference's own documented output (KAT-1, doc/doc-egg-gencat.tex:32-55: this pins the C++ FITS
reader and the reference-wavelength / FWHM rules of src/egg-gencat.cpp:195-199), user errors with
exit code 1 (src/egg-gencat.cpp:286-289), and the loud failure without a GPU.
GPU tier: a full run from FITS inputs to the FITS output column groups and the save_sed files."""
import os
import subprocess

import numpy as np
import pytest

import cases
from egg_b200 import fitscol, synth
from test_oracle_kat import KAT1

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "egg_b200", "egg-gencat-b200")


def _need_cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "egg_b200", "csrc"), "-s", "all"])


def _filter_db(tmp_path, names):
    """A filter data base in the reference's layout (db.dat + one-row LAM/RES tables) from the packed curves."""
    pack = synth.load_pack()
    d = tmp_path / "filter-db"
    (d / "f").mkdir(parents=True, exist_ok=True)
    lines = []
    for n in names:
        lam, res = pack[f"filter/{n}/lam"], pack[f"filter/{n}/res"]  # original dtypes: four curves are f32 ('E')
        fitscol.write_table(str(d / "f" / f"{n}.fits"), {"LAM": lam, "RES": res})
        lines.append(f"{n}=f/{n}.fits")
    (d / "db.dat").write_text("\n".join(lines) + "\n")
    return str(d / "db.dat")


def _run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def test_list_bands_reproduces_the_reference_manual(tmp_path):
    _need_cli()
    db = _filter_db(tmp_path, sorted(KAT1) + ["hst-f160w"])
    for pattern, want in (("jwst-", [n for n in sorted(KAT1) if n.startswith("jwst-")]),
                          ("-Ks", [n for n in sorted(KAT1) if n.endswith("-Ks")])):
        r = _run(f"list_bands={pattern}", f"filter_db={db}")
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        assert lines[0] == f"List of available bands (filter: '{pattern}'):"
        got = {}
        for ln in lines[1:]:
            name = ln.split()[1]
            got[name] = (ln.split("ref-lam = ")[1].split(" um")[0], ln.split("FWHM = ")[1].split(" um")[0])
        assert sorted(got) == want
        for n in want:  # the digits printed in doc/doc-egg-gencat.tex:32-55
            assert got[n] == (f"{KAT1[n][0]:g}", f"{KAT1[n][1]:g}"), (n, got[n])
        # sorted by instrument, then by reference wavelength (src/egg-gencat.cpp:208-215)
        order = [ln.split()[1] for ln in lines[1:]]
        key = [(n.split("-")[0], KAT1[n][0]) for n in order]
        assert key == sorted(key)


def test_user_errors_exit_1_like_the_reference(tmp_path):
    _need_cli()
    r = _run("igm=foo")
    assert r.returncode == 1 and "unknown IGM prescription 'foo'" in r.stderr
    r = _run("area=0.08", "mmin=9")
    assert r.returncode == 1 and "input_props" in r.stderr          # catalog generation is out of scope
    r = _run("list_bands", f"filter_db={tmp_path}/nope.dat")
    assert r.returncode == 1 and "could not find filter database file" in r.stderr
    assert _run("help").returncode == 0


def _write_inputs(tmp_path, ngal=300, bands=synth.B4 + ["spitzer-irac1", "herschel-pacs100"], rf=synth.RF3):
    L = cases.libs()
    opt, ir = L["opt"], L["ce01"]
    gal = synth.draw_galaxies(ngal, opt_lib=opt, ir_lib=ir, mmin=7.5, seed=5)
    db = _filter_db(tmp_path, bands + rf)
    fitscol.write_table(str(tmp_path / "opt_lib.fits"), {k.upper(): np.asarray(v) for k, v in opt.items()})
    fitscol.write_table(str(tmp_path / "ir_lib_ce01.fits"), {"LAM": np.asarray(ir["lam"]), "SED": np.asarray(ir["sed"])})
    cat = {"ID": np.arange(ngal, dtype=np.uint64), "RA": np.zeros(ngal), "DEC": np.zeros(ngal)}
    for k in ("z", "d", "m_disk", "m_bulge", "m2lcor", "lir", "igm_abs", "vdisp", "line_lum_disk", "line_lum_bulge",
              "opt_sed_disk", "opt_sed_bulge", "ir_sed"):
        cat[k.upper()] = np.asarray(gal[k])
    cat["LINES"] = np.array(synth.LINES)
    cat["LINE_LAMBDA"] = synth.LINE_LAMBDA
    fitscol.write_table(str(tmp_path / "cat.fits"), cat)
    args = [f"input_props={tmp_path}/cat.fits", f"filter_db={db}", f"opt_lib={tmp_path}/opt_lib.fits",
            f"ir_lib={tmp_path}/ir_lib_ce01.fits", "bands=[" + ",".join(bands) + "]", "rfbands=[" + ",".join(rf) + "]"]
    return opt, ir, gal, bands, rf, args


def test_without_a_gpu_the_stage_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _need_cli()
    *_, args = _write_inputs(tmp_path, ngal=8)
    r = _run(*args, f"out={tmp_path}/out.fits")
    assert r.returncode == 1 and "CUDA" in r.stderr and not os.path.exists(tmp_path / "out.fits")


@pytest.mark.gpu
def test_full_run_writes_the_reference_column_groups(tmp_path):
    import egg_b200
    from oracle.oracle import Oracle
    _need_cli()
    opt, ir, gal, bands, rf, args = _write_inputs(tmp_path)
    out = str(tmp_path / "out.fits")
    r = _run(*args, f"out={out}", "save_sed", "verbose")
    assert r.returncode == 0, r.stderr
    t = fitscol.read_table(out)
    names = list(t)
    for grp in (["FLUX", "FLUX_DISK", "FLUX_BULGE", "BANDS", "LAMBDA"], ["RFMAG", "RFMAG_DISK", "RFMAG_BULGE", "RFBANDS", "RFLAMBDA"]):
        assert [n for n in names if n in grp] == grp                 # src/egg-gencat.cpp:2266-2286
    assert t["Z"].shape == (300,) and t["FLUX"].shape == (300, len(bands)) and t["FLUX"].dtype == np.float32
    ph = egg_b200.Photometry(opt, ir, synth.get_filters(bands), synth.get_filters(rf), line_lambda=synth.LINE_LAMBDA)
    py = ph.compute(gal)
    assert list(t["BANDS"]) == [bands[k] for k in ph.band_order]     # sorted by reference wavelength (:367-378)
    np.testing.assert_array_equal(t["LAMBDA"], ph.lambda_)
    for k in ("FLUX", "FLUX_DISK", "FLUX_BULGE", "RFMAG", "RFMAG_DISK", "RFMAG_BULGE"):
        np.testing.assert_array_equal(t[k], py[k.lower()])           # same library underneath: identical
    ref = Oracle(opt, ir, synth.get_filters(bands), synth.get_filters(rf), line_lambda=synth.LINE_LAMBDA).compute(gal, f64=True)
    assert cases.rel_err(t["FLUX_DISK"], ref["flux_disk"], 1e-5).max() < 2e-5
    # save_sed: all bulges then all disks, lam then sed as f32 (src/egg-gencat.cpp:2164-2193, 2230-2235)
    lk = fitscol.read_table(str(tmp_path / "out-seds-lookup.fits"))
    raw = np.fromfile(str(tmp_path / "out-seds.dat"), dtype=np.float32)
    nb, nd = ph.sed_size("bulge"), ph.sed_size("disk")
    assert int(lk["ELEM_SIZE"].ravel()[0]) == 4 and raw.size == 300 * 2 * (nb + nd)
    assert lk["BULGE_START"][0] == 0 and lk["DISK_START"][0] == 300 * 8 * nb and lk["DISK_NBYTE"][7] == 8 * nd
    lam, sed = ph.get_seds(gal, 7, 1, "disk")
    o = int(lk["DISK_START"][7]) // 4
    np.testing.assert_array_equal(raw[o:o + nd], lam[0])
    np.testing.assert_array_equal(raw[o + nd:o + 2 * nd], sed[0])
    ph.close()
    # two shards on one device count as two "GPUs" only where two exist; one GPU: gpus=1 is the same file
    r = _run(*args, f"out={tmp_path}/fp64.fits", "precision=fp64")
    assert r.returncode == 0, r.stderr
    t64 = fitscol.read_table(str(tmp_path / "fp64.fits"))
    assert cases.rel_err(t64["FLUX_DISK"], ref["flux_disk"], 1e-6).max() < 2e-7   # f32 column of the fp64 result


@pytest.mark.gpu
def test_two_gpu_shards_give_the_same_catalog(tmp_path):
    """gpus=2: two contiguous shards, one context and host thread per GPU, no exchange."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _need_cli()
    *_, args = _write_inputs(tmp_path, ngal=3001)
    assert _run(*args, f"out={tmp_path}/one.fits").returncode == 0
    r = _run(*args, f"out={tmp_path}/two.fits", "gpus=2")
    assert r.returncode == 0, r.stderr
    a, b = fitscol.read_table(str(tmp_path / "one.fits")), fitscol.read_table(str(tmp_path / "two.fits"))
    for k in ("FLUX", "FLUX_DISK", "FLUX_BULGE", "RFMAG", "RFMAG_DISK", "RFMAG_BULGE"):
        np.testing.assert_array_equal(a[k], b[k])


@pytest.mark.gpu
def test_input_cat_generates_the_properties_on_the_gpu(tmp_path):
    """input_cat=: ID RA DEC Z M PASSIVE in, every property column (src/egg-gencat.cpp:1360-2022) and the fluxes out."""
    import egg_b200
    from egg_b200 import properties
    _need_cli()
    L = cases.libs()
    opt, ir = L["opt"], L["ce01"]
    n = 400
    rng = np.random.default_rng(12)
    z, m = rng.uniform(0.1, 6, n).astype(np.float32), rng.uniform(8, 11.5, n).astype(np.float32)
    pas = rng.random(n) < 0.3
    bands = synth.B4 + ["spitzer-irac1"]
    db = _filter_db(tmp_path, bands)
    fitscol.write_table(str(tmp_path / "opt_lib.fits"), {k.upper(): np.asarray(v) for k, v in opt.items()})
    fitscol.write_table(str(tmp_path / "ir_lib_ce01.fits"), {"LAM": np.asarray(ir["lam"]), "SED": np.asarray(ir["sed"])})
    fitscol.write_table(str(tmp_path / "zm.fits"), {"ID": np.arange(n, dtype=np.uint64), "RA": np.zeros(n), "DEC": np.zeros(n),
                                                     "Z": z, "M": m, "PASSIVE": pas})
    args = [f"input_cat={tmp_path}/zm.fits", f"filter_db={db}", f"opt_lib={tmp_path}/opt_lib.fits", f"ir_lib={tmp_path}/ir_lib_ce01.fits",
            "bands=[" + ",".join(bands) + "]", f"out={tmp_path}/out.fits", "seed=7"]
    r = _run(*args)
    assert r.returncode == 0, r.stderr
    t = fitscol.read_table(str(tmp_path / "out.fits"))
    pr = properties.Properties(opt, seed=7)
    want = pr.compute(z, m, pas)
    pr.close()
    for k, v in want.items():
        np.testing.assert_array_equal(t[k.upper()], v, err_msg=k)                       # same library underneath
    assert list(t["LINES"]) == properties.line_table()[0]
    np.testing.assert_array_equal(t["LINE_LUM"], want["line_lum_disk"] + want["line_lum_bulge"])
    np.testing.assert_array_equal(t["PASSIVE"], pas)                                   # input columns pass through
    gal = {k: want[k] for k in ("d", "m_disk", "m_bulge", "m2lcor", "opt_sed_disk", "opt_sed_bulge", "ir_sed", "lir", "igm_abs",
                                "vdisp", "line_lum_disk", "line_lum_bulge")}
    gal["z"] = z
    ph = egg_b200.Photometry(opt, ir, synth.get_filters(bands), line_lambda=properties.line_table()[1])
    f = ph.compute(gal)
    ph.close()
    np.testing.assert_array_equal(t["FLUX"], f["flux"])
    # the column contract of the output catalog (src/egg-gencat.cpp:2253-2286): the writer's order, and the lists the
    # downstream programs ask for by name (src/egg-2skymaker.cpp:99-104, src/egg-genmap.cpp:94-102), with their types
    ref_order = ["ID", "RA", "DEC", "CLUSTERED", "Z", "D", "M", "SFR", "RSB", "M2LCOR", "DISK_ANGLE", "DISK_RADIUS", "DISK_RATIO",
                 "BULGE_ANGLE", "BULGE_RADIUS", "BULGE_RATIO", "BT", "M_DISK", "M_BULGE", "AV_DISK", "AV_BULGE", "PASSIVE",
                 "RFUV_BULGE", "RFUV_DISK", "RFVJ_BULGE", "RFVJ_DISK", "SFRIR", "SFRUV", "IRX", "LIR", "IR_SED",
                 "OPT_SED_BULGE", "OPT_SED_DISK", "TDUST", "IR8", "FPAH", "MDUST", "IGM_ABS", "ZB", "CMD",
                 "FLUX", "FLUX_DISK", "FLUX_BULGE", "BANDS", "LAMBDA",
                 "VDISP", "AVLINES_DISK", "AVLINES_BULGE", "LINE_LUM", "LINE_LUM_DISK", "LINE_LUM_BULGE", "LINES", "LINE_LAMBDA"]
    names = list(t)
    assert [c for c in names if c in ref_order] == [c for c in ref_order if c in names]
    assert set(ref_order) - set(names) <= {"ZB", "CMD"}      # (written by the generation stage only: redshift bins, command line)
    for c in ("ID", "RA", "DEC", "DISK_ANGLE", "DISK_RADIUS", "DISK_RATIO", "BULGE_ANGLE", "BULGE_RADIUS", "BULGE_RATIO",
              "FLUX_DISK", "FLUX_BULGE", "BANDS", "LAMBDA"):   # egg-2skymaker
        assert c in t, c
    for c in ("RA", "DEC", "FLUX", "BANDS"):                   # egg-genmap
        assert c in t, c
    for c in ("ID", "OPT_SED_DISK", "OPT_SED_BULGE", "IR_SED"):
        assert t[c].dtype.kind in "iu" and t[c].dtype.itemsize == 8, c                                              # TFORM K
    assert t["RA"].dtype == np.float64 and t["DEC"].dtype == np.float64                                             # D
    assert t["CLUSTERED"].dtype == bool and t["PASSIVE"].dtype == bool and not t["CLUSTERED"].any()                  # L (:1356)
    for c in ("Z", "D", "M", "FLUX", "FLUX_DISK", "FLUX_BULGE", "LAMBDA", "IGM_ABS", "LINE_LUM"):
        assert t[c].dtype == np.float32, c                                                                           # E
    assert t["FLUX"].shape == (n, len(bands)) and t["IGM_ABS"].shape == (n, 3) and len(t["BANDS"]) == len(bands)
    # the reference's input check (src/egg-gencat.cpp:769-774)
    fitscol.write_table(str(tmp_path / "bad.fits"), {"ID": np.arange(2, dtype=np.uint64), "Z": np.array([0.5, -1.0], np.float32),
                                                      "M": np.array([10.0, 10.0], np.float32), "PASSIVE": np.array([False, True])})
    r = _run(*[x.replace("zm.fits", "bad.fits") for x in args])
    assert r.returncode == 1 and "1 unphysical galaxy" in r.stderr


def test_input_cat_is_validated_before_the_device_is_touched(tmp_path):
    """CPU tier: the reference's input checks (src/egg-gencat.cpp:714-779) and the loud failure without a GPU."""
    _need_cli()
    L = cases.libs()
    fitscol.write_table(str(tmp_path / "opt_lib.fits"), {k.upper(): np.asarray(v) for k, v in L["opt"].items()})
    fitscol.write_table(str(tmp_path / "ir_lib_ce01.fits"), {"LAM": np.asarray(L["ce01"]["lam"]), "SED": np.asarray(L["ce01"]["sed"])})
    db = _filter_db(tmp_path, synth.B4)
    common = [f"filter_db={db}", f"opt_lib={tmp_path}/opt_lib.fits", f"ir_lib={tmp_path}/ir_lib_ce01.fits",
              "bands=[" + ",".join(synth.B4) + "]", f"out={tmp_path}/out.fits"]
    ok = {"ID": np.arange(3, dtype=np.uint64), "Z": np.array([0.5, 1.0, 2.0], np.float32), "M": np.array([9.0, 10.0, 11.0], np.float32),
          "PASSIVE": np.array([False, True, False])}
    fitscol.write_table(str(tmp_path / "nom.fits"), {k: v for k, v in ok.items() if k != "M"})
    r = _run(f"input_cat={tmp_path}/nom.fits", *common)
    assert r.returncode == 1 and "no column M" in r.stderr
    bad = dict(ok, M=np.array([9.0, 3.5, 14.0], np.float32))
    fitscol.write_table(str(tmp_path / "bad.fits"), bad)
    r = _run(f"input_cat={tmp_path}/bad.fits", *common)
    assert r.returncode == 1 and "2 unphysical galaxies" in r.stderr and "4 < m < 13" in r.stdout + r.stderr
    fitscol.write_table(str(tmp_path / "lir.fits"), dict(ok, LIR=np.ones(3, np.float32)))
    r = _run(f"input_cat={tmp_path}/lir.fits", *common)
    assert r.returncode == 1 and "LIR" in r.stderr
    fitscol.write_table(str(tmp_path / "ok.fits"), ok)
    r = _run(f"input_cat={tmp_path}/ok.fits", *[c.replace("ir_lib_ce01.fits", "ir_lib_other.fits") for c in common])
    assert r.returncode == 1                                                           # unreadable / uncalibrated IR library
    import torch
    if not torch.cuda.is_available():
        r = _run(f"input_cat={tmp_path}/ok.fits", *common)
        assert r.returncode == 1 and "CUDA" in r.stderr and not os.path.exists(tmp_path / "out.fits")


def test_ascii_input_catalog_is_read_like_the_reference(tmp_path):
    """Plain-text input_cat (ID RA DEC Z M PASSIVE, src/egg-gencat.cpp:765-767): same checks as the FITS form."""
    _need_cli()
    L = cases.libs()
    fitscol.write_table(str(tmp_path / "opt_lib.fits"), {k.upper(): np.asarray(v) for k, v in L["opt"].items()})
    fitscol.write_table(str(tmp_path / "ir_lib_ce01.fits"), {"LAM": np.asarray(L["ce01"]["lam"]), "SED": np.asarray(L["ce01"]["sed"])})
    db = _filter_db(tmp_path, synth.B4)
    common = [f"filter_db={db}", f"opt_lib={tmp_path}/opt_lib.fits", f"ir_lib={tmp_path}/ir_lib_ce01.fits",
              "bands=[" + ",".join(synth.B4) + "]", f"out={tmp_path}/out.fits"]
    (tmp_path / "bad.txt").write_text("# id ra dec z m passive\n1 53.1 -27.8 0.5 10.2 0\n2 53.1 -27.8 -0.1 10.0 1\n")
    r = _run(f"input_cat={tmp_path}/bad.txt", *common)
    assert r.returncode == 1 and "1 unphysical galaxy" in r.stderr
    (tmp_path / "short.txt").write_text("1 53.1 -27.8 0.5\n")
    r = _run(f"input_cat={tmp_path}/short.txt", *common)
    assert r.returncode == 1 and "expected ID RA DEC Z M PASSIVE" in r.stderr
    (tmp_path / "ok.txt").write_text("# id ra dec z m passive\n1 53.1 -27.8 0.5 10.2 0\n2 53.2 -27.9 1.5 11.0 1\n")
    import torch
    r = _run(f"input_cat={tmp_path}/ok.txt", *common, "seed=3")
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "CUDA" in r.stderr                     # read and validated, then no device
    else:
        assert r.returncode == 0, r.stderr
        t = fitscol.read_table(str(tmp_path / "out.fits"))
        assert list(t["ID"]) == [1, 2] and list(t["PASSIVE"]) == [False, True] and t["FLUX"].shape == (2, 4)


def test_maglim_needs_a_selection_band(tmp_path):
    """src/egg-gencat.cpp:240-243: a magnitude limit without a selection band is a user error (exit 1, no device touched)."""
    _need_cli()
    r = _run("maglim=26", f"filter_db={_filter_db(tmp_path, ['hst-f160w'])}")
    assert r.returncode == 1 and "selection band" in r.stderr


@pytest.mark.gpu
def test_maglim_prints_the_mass_limit_per_redshift_slice(tmp_path):
    """`maglim= selection_band=` without a catalog: z_mmin per slice of make_zbins (src/egg-gencat.cpp:803-839, 903-969),
    the same numbers as the library entry point on the same slices."""
    import egg_b200
    from egg_b200 import properties
    _need_cli()
    opt = cases.libs()["opt"]
    db = _filter_db(tmp_path, ["hst-f160w"])
    fitscol.write_table(str(tmp_path / "opt_lib.fits"), {k.upper(): np.asarray(v) for k, v in opt.items()})
    r = _run("maglim=26", "selection_band=hst-f160w", f"filter_db={db}", f"opt_lib={tmp_path}/opt_lib.fits", "zmin=0.1", "zmax=6", "dz=0.3",
             "mmin_strict=6", "verbose")
    assert r.returncode == 0, r.stderr
    rows = np.array([[float(v) for v in line.split()] for line in r.stdout.splitlines() if line[:1].isdigit()])
    lo, hi, zm = rows[:, 0], rows[:, 1], rows[:, 2]
    assert lo[0] == pytest.approx(0.1) and hi[-1] == pytest.approx(6.0) and np.all(lo[1:] == hi[:-1])
    assert np.all(np.diff(lo) <= 0.5 + 1e-9)                          # max_dz
    assert "will generate masses from as low as" in r.stdout + r.stderr
    ph = egg_b200.Photometry(opt, cases.libs()["ce01"], synth.get_filters(["hst-f160w"]), no_dust=True, no_nebular=True, igm="none")
    pr = properties.Properties(opt, no_random=True)
    want = ph.maglim(pr, lo, hi, 26.0, 6.0, 12.0, 1000)
    pr.close()
    ph.close()
    np.testing.assert_allclose(zm, want, rtol=0, atol=2e-5)            # (printed with six significant digits)


@pytest.mark.gpu
def test_check_mode_compares_with_the_fluxes_already_in_the_catalog(tmp_path):
    """an egg output catalog holds every input column of this stage next to FLUX* / RFMAG*, so it is a
    test vector.  `check` recomputes and compares per band: exit 0 within tolerance, 2 above it; `sed2flux_nodes=` selects the
    convention (a catalog written under another node set fails under `merged` and passes under its own)."""
    from oracle.oracle import Oracle
    _need_cli()
    opt, ir, gal, bands, rf, args = _write_inputs(tmp_path, ngal=200)
    fb, fr = synth.get_filters(bands), synth.get_filters(rf)
    cat = dict(fitscol.read_table(str(tmp_path / "cat.fits")))

    def write(name, nodes, spoil=None):
        O = Oracle(opt, ir, fb, fr, line_lambda=synth.LINE_LAMBDA, sed2flux_nodes=nodes)
        ref = O.compute(gal)
        c = dict(cat)
        for k in ("flux", "flux_disk", "flux_bulge", "rfmag", "rfmag_disk", "rfmag_bulge"):
            c[k.upper()] = np.asarray(ref[k], np.float32)
        c["BANDS"] = np.array([bands[k] for k in O.band_order])
        c["RFBANDS"] = np.array([rf[k] for k in O.rfband_order])
        if spoil is not None:
            c["FLUX_DISK"] = c["FLUX_DISK"].copy()
            c["FLUX_DISK"][:, spoil] *= np.float32(1.001)
        fitscol.write_table(str(tmp_path / name), c)
        return [x.replace("cat.fits", name) for x in args if not x.startswith(("bands=", "rfbands="))]   # bands from the catalog itself

    ok = write("egg_merged.fits", "merged")
    r = _run(*ok, "check", "check_tol=2e-5")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "every compared band is within tolerance" in r.stdout and r.stdout.count("check: FLUX_DISK") == len(bands)
    assert not any(p.name.startswith("egg-2") for p in tmp_path.iterdir())          # nothing is written without out=
    bad = write("egg_spoiled.fits", "merged", spoil=2)
    r = _run(*bad, "check", "check_tol=2e-5")
    assert r.returncode == 2 and "DIFFERENCES above tolerance" in r.stdout
    lines = [l for l in r.stdout.splitlines() if l.startswith("check: FLUX_DISK")]
    assert sum("above 2e-05: 0 of" not in l for l in lines) == 1                     # exactly the spoiled band
    other = write("egg_sed_nodes.fits", "sed")
    assert _run(*other, "check", "check_tol=2e-5").returncode == 2                  # the node sets differ by more than the tolerance
    r = _run(*other, "check", "check_tol=2e-5", "sed2flux_nodes=sed")
    assert r.returncode == 0, r.stdout + r.stderr
