"""CPU tier: the C-ABI library loads, exports every symbol include/eggphot.h declares, validates its
arguments before touching the device, and fails loudly (no CPU path) when no GPU is present."""
import os
import re

import numpy as np
import pytest

import cases
import egg_b200
from egg_b200 import synth
from egg_b200.photometry import EXPORTS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "eggphot.h")).read()
    declared = set(re.findall(r"\b(eggphot_[a-z_]+)\s*\(", hdr))
    lib = egg_b200.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(EXPORTS)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_user_errors_are_reported_before_the_device_is_touched():
    L = cases.libs()
    bands = synth.get_filters(synth.B4)
    with pytest.raises(egg_b200.EggphotError) as e:
        egg_b200.Photometry(L["opt"], L["ce01"], bands, igm="bogus")
    assert e.value.code == 1 and "unknown IGM prescription" in str(e.value)          # gencat:286-289
    with pytest.raises(egg_b200.EggphotError) as e:
        egg_b200.Photometry(L["opt"], L["ce01"], bands, opt_lib_imf="scalo")
    assert e.value.code == 3 and "unknown IMF" in str(e.value)                        # gencat:464
    unusable = dict(L["opt"], use=np.zeros_like(L["opt"]["use"]))
    with pytest.raises(egg_b200.EggphotError) as e:
        egg_b200.Photometry(unusable, L["ce01"], bands)
    assert e.value.code == 3 and "does not contain any valid SED" in str(e.value)     # gencat:537-541
    with pytest.raises(egg_b200.EggphotError) as e:
        egg_b200.Photometry(L["opt"], L["ce01"], bands, no_stellar=True)             # ce01 starts at 0.1 um
    assert e.value.code == 4
    with pytest.raises(egg_b200.EggphotError) as e:
        egg_b200.Photometry(L["opt"], L["ce01"], bands * 70)                          # more band tables than the pipeline holds
    assert e.value.code == 5
    bad = [(np.array([1.0, 0.9, 1.1]), np.ones(3))]
    with pytest.raises(egg_b200.EggphotError) as e:
        egg_b200.Photometry(L["opt"], L["ce01"], bad)
    assert e.value.code == 2


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_an_error_not_a_cpu_fallback():
    L = cases.libs()
    with pytest.raises(egg_b200.EggphotError) as e:
        egg_b200.Photometry(L["opt"], L["ce01"], synth.get_filters(synth.B4))
    assert e.value.code == 7 and "no CPU path" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "egg_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("oracle restatement", "").lower() or f == "synth.py", f
