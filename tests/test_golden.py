"""Golden vectors (tests/golden/*.npz, made by tools/make_golden.py from the oracle): the oracle must
keep reproducing them bit for bit, and the device algorithm must match them within tolerance."""
import os

import numpy as np
import pytest

import cases
import emul
from oracle.oracle import Oracle

GOLDEN = ("b4_defaults", "b30_rf3_defaults", "b40_hd_cs17", "b23_no_dust", "b30_hd_noigm")  # (the last: BASELINE configs[4])
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    c = cases.make(name, ngal=256, seed=2024)
    g = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    chk = np.array([np.float64(c["gal"]["z"].sum()), np.float64(c["gal"]["lir"].astype(np.float64).sum()),
                    np.float64(c["opt"]["sed"].astype(np.float64).sum())])
    np.testing.assert_allclose(chk, g["input_checksum"], rtol=1e-12, err_msg="synthetic inputs changed: regenerate the golden files")
    return c, g


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_reproduces_golden(name):
    c, g = load(name)
    ref = Oracle(c["opt"], c["ir"], c["bands"], c["rfbands"], line_lambda=c["lines"], **c["opts"]).compute(c["gal"], f64=True)
    for k in ("flux_disk", "flux_bulge", "rfmag_disk", "rfmag_bulge"):
        np.testing.assert_allclose(ref[k], g[k], rtol=1e-13, atol=0)


@pytest.mark.parametrize("name", GOLDEN[:2])
def test_device_algorithm_matches_golden(name):
    c, g = load(name)
    out = emul.compute(c["opt"], c["ir"], c["bands"], c["rfbands"], c["gal"], line_lambda=c["lines"], precision="fp32c", **c["opts"])
    for k in ("flux_disk", "flux_bulge"):
        assert cases.rel_err(out[k], g[k], cases.FLOOR["fp32c"]).max() < cases.TOL["fp32c"]
