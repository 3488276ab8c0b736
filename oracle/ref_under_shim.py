"""Import the UNMODIFIED reference package from /root/reference on top of the
NumPy JAX shim (oracle/jax_shim).  TEST INFRASTRUCTURE ONLY: used by
tests/golden/make_golden.py in the build container to generate golden vectors;
/root/reference does not exist on the GPU box, so nothing at run time depends
on this module (tests that would use it skip when the path is absent).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("METMHN_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jax_shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "metmhn", "jx"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    colors = types.ModuleType("matplotlib.colors")
    plt.Axes = object
    colors.LinearSegmentedColormap = object
    mpl.pyplot, mpl.colors = plt, colors
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.colors": colors})


def load():
    """Return (regularized_optimization, likelihood, one_event, vanilla, kronvec) of the reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__file__", "").startswith(_SHIM):
        raise RuntimeError("a real jax is already imported; the shim is not needed")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    _stub_matplotlib()
    regopt = importlib.import_module("metmhn.regularized_optimization")
    lik = importlib.import_module("metmhn.jx.likelihood")
    one = importlib.import_module("metmhn.jx.one_event")
    van = importlib.import_module("metmhn.jx.vanilla")
    kv = importlib.import_module("metmhn.jx.kronvec")
    return regopt, lik, one, van, kv
