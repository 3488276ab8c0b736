import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_err(a, b, floor=1e-3):
    """Entrywise relative error with the tolerance semantics of DESIGN.md section 6:
    |a-b| / max(|b|, floor*max|b|)."""
    import numpy as np
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), floor * np.max(np.abs(b)) if b.size else 0.0)
    scale = np.where(scale == 0.0, 1.0, scale)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
