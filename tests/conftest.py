import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
np.seterr(all="ignore")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, as __graft_entry__.build() does
    need = [os.path.join(ROOT, "egg_b200", "libeggphot.so"), os.path.join(ROOT, "egg_b200", "egg-gencat-b200"),
            os.path.join(ROOT, "oracle", "libeggphot_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def libs():
    from egg_b200 import synth
    return {"opt": synth.make_opt_lib(), "ce01": synth.ce01_lib(), "cs17": synth.make_cs17_lib(),
            "opt_hd": None}
