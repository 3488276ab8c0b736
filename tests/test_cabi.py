"""CPU tests of the boundary: the shared library builds, loads and exports every symbol of
include/metmhn_b200.h; compute calls fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "metmhn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmh_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from metmhn_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert set(_lib.EXPORTS) == set(names)
    for name in names:
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from metmhn_b200 import Handle, MetMHNError
    with pytest.raises(MetMHNError) as e:
        Handle(np.zeros((2, 9), dtype=np.int8))
    assert e.value.code == -2


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "metmhn_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_penalties_match_oracle():
    from metmhn_b200 import regularized_optimization as ro
    from oracle import reference_restated as rr
    rng = np.random.default_rng(0)
    p = rng.normal(size=6 * 8)
    v, g = ro.symmetric_penal(p, 6)
    v0, g0 = rr.symmetric_penal(p, 6)
    assert abs(v - v0) < 1e-14 and np.abs(g - g0).max() < 1e-14
