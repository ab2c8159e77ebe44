import numpy as np

from egg_b200 import fitscol


def test_roundtrip_column_oriented_table(tmp_path):
    cols = {"ID": np.arange(5, dtype=np.uint64), "RA": np.linspace(0, 1, 5), "Z": np.arange(5, dtype=np.float32),
            "FLUX": np.arange(15, dtype=np.float32).reshape(5, 3), "PASSIVE": np.array([True, False, True, True, False]),
            "BANDS": np.array(["hst-f160w", "a", "spitzer-irac1"]), "CMD": "egg-gencat area=0.08"}
    p = str(tmp_path / "t.fits")
    fitscol.write_table(p, cols)
    t = fitscol.read_table(p)
    assert list(t) == list(cols)
    np.testing.assert_array_equal(t["ID"], cols["ID"].astype(np.int64))
    np.testing.assert_array_equal(t["FLUX"], cols["FLUX"])
    assert t["FLUX"].shape == (5, 3)              # TDIM = '(3,5)': fastest axis first
    np.testing.assert_array_equal(t["PASSIVE"], cols["PASSIVE"])
    assert list(t["BANDS"]) == list(cols["BANDS"])
    assert str(t["CMD"]) == cols["CMD"]
    raw = open(p, "rb").read()
    assert len(raw) % 2880 == 0 and b"TDIM4   = '(3,5)" in raw


def test_filter_db_parser(tmp_path):
    (tmp_path / "db.dat").write_text("a-b=x/y.fits\n# comment\n\nc-d=z.fits\n")
    db = fitscol.read_filter_db(str(tmp_path / "db.dat"))
    assert list(db) == ["a-b", "c-d"] and db["a-b"].endswith("x/y.fits")
