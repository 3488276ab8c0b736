"""jax.numpy stand-in: NumPy FP64 with the functional `.at[]` update helper.
TEST INFRASTRUCTURE ONLY (see package docstring)."""
from __future__ import annotations

import numpy as _np

int8 = _np.int8
int32 = _np.int32
int64 = _np.int64
float64 = _np.float64


class _AtIndex:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def _idx_np(self):
        idx = self._idx
        # jnp accepts tuples of column indices where NumPy wants lists
        if isinstance(idx, tuple):
            idx = tuple(list(i) if isinstance(i, tuple) else i for i in idx)
        return idx

    def set(self, value):
        out = _np.array(self._arr, copy=True)
        out[self._idx_np()] = value
        return _wrap(out)

    def add(self, value):
        out = _np.array(self._arr, copy=True)
        _np.add.at(out, self._idx_np(), value)
        return _wrap(out)

    def multiply(self, value):
        out = _np.array(self._arr, copy=True)
        out[self._idx_np()] = out[self._idx_np()] * value
        return _wrap(out)

    def divide(self, value):
        out = _np.array(self._arr, copy=True)
        out[self._idx_np()] = out[self._idx_np()] / value
        return _wrap(out)

    def get(self):
        return _wrap(_np.asarray(self._arr)[self._idx_np()])


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self._arr, idx)


class ndarray(_np.ndarray):
    """ndarray subclass that carries `.at`; everything else is plain NumPy."""

    @property
    def at(self):
        return _At(self)


def _wrap(x):
    a = _np.asarray(x)
    if a.dtype == _np.float32:
        a = a.astype(_np.float64)
    return a.view(ndarray)


def array(x, dtype=None):
    return _wrap(_np.array(x, dtype=dtype))


def asarray(x, dtype=None):
    return _wrap(_np.asarray(x, dtype=dtype))


def _lift(f):
    def g(*a, **k):
        r = f(*a, **k)
        if isinstance(r, _np.ndarray):
            return _wrap(r)
        if isinstance(r, tuple):
            return tuple(_wrap(t) if isinstance(t, _np.ndarray) else t for t in r)
        return r
    g.__name__ = getattr(f, "__name__", "lifted")
    return g


for _name in ("zeros", "ones", "zeros_like", "ones_like", "exp", "log", "log2", "sqrt", "sum",
              "dot", "diag", "diagonal", "diag_indices", "apply_along_axis", "arange", "append",
              "min", "max", "hstack", "vstack", "unique", "concatenate", "ceil", "floor", "abs",
              "stack", "cumsum", "prod", "eye", "outer", "matmul", "transpose", "row_stack",
              "all", "any", "isclose", "allclose", "linspace", "argmax", "argmin"):
    if hasattr(_np, _name):
        globals()[_name] = _lift(getattr(_np, _name))
if "row_stack" not in globals():
    row_stack = _lift(_np.vstack)


def where(cond, x=None, y=None, size=None, fill_value=0):
    if x is not None or y is not None:
        return _wrap(_np.where(cond, x, y))
    idx = _np.where(_np.asarray(cond))
    if size is not None:
        out = []
        for ix in idx:
            ix = ix[:size]
            if ix.shape[0] < size:
                ix = _np.concatenate([ix, _np.full(size - ix.shape[0], fill_value, dtype=ix.dtype)])
            out.append(_wrap(ix))
        return tuple(out)
    return tuple(_wrap(ix) for ix in idx)
