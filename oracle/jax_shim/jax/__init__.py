"""Minimal NumPy stand-in for the subset of the JAX API that cbg-ethz/metMHN uses.

TEST INFRASTRUCTURE ONLY.  JAX/jaxlib are not installable in this image (no
network, no wheel), so the reference package under /root/reference cannot run
on its real backend.  This shim lets the UNMODIFIED reference sources be
imported and executed with NumPy FP64 arrays standing in for XLA buffers, so
that `tests/golden/make_golden.py` can produce golden vectors from the
reference's own code.  All of the reference's arithmetic on this path is
element-wise FP64, exp/log, sums and 2x2/4x4 matmuls, whose NumPy semantics
(`reshape(order="C")`, `flatten(order="F")`, `@`) are identical to jnp's.

Only what the hot path touches is provided: jit (identity), vmap (Python
loop + stack), lax.{fori_loop,cond,switch,select_n,select,dynamic_slice,
while_loop}, the `.at[...]` functional-update helper, and a few jnp
functions.  Nothing in the product imports this package.
"""
from __future__ import annotations

import functools

import numpy as _np

from . import numpy as numpy  # noqa: F401  (jax.numpy)
from . import lax as lax  # noqa: F401
from . import random as random  # noqa: F401
from .numpy import _wrap


class _Config:
    def update(self, *_a, **_k):
        return None


config = _Config()


def jit(fun=None, **_kw):
    """Identity: there is nothing to trace."""
    if fun is None:
        return lambda f: f
    return fun


def _take(arg, axis, i):
    if axis is None:
        return arg
    return _wrap(_np.take(_np.asarray(arg), i, axis=axis))


def vmap(fun, in_axes=0, out_axes=0):
    """Map `fun` over the leading axis by a Python loop and stack the results."""

    @functools.wraps(fun)
    def mapped(*args):
        axes = in_axes
        if not isinstance(axes, (tuple, list)):
            axes = (axes,) * len(args)
        length = None
        for a, ax in zip(args, axes):
            if ax is not None:
                length = _np.asarray(a).shape[ax]
                break
        outs = [fun(*[_take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(length)]
        if isinstance(outs[0], tuple):
            return tuple(_wrap(_np.stack([_np.asarray(o[j]) for o in outs], axis=0))
                         for j in range(len(outs[0])))
        return _wrap(_np.stack([_np.asarray(o) for o in outs], axis=0))

    return mapped
