"""jax.lax stand-in: eager Python control flow.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as _np

from .numpy import _wrap


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def cond(pred, true_fun, false_fun, *operands, operand=None):
    if operand is not None:
        operands = (operand,)
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def switch(index, branches, *operands, operand=None):
    if operand is not None:
        operands = (operand,)
    i = min(max(int(index), 0), len(branches) - 1)   # XLA clamps the branch index
    return branches[i](*operands)


def select_n(which, *cases):
    return cases[int(which)]


def select(pred, on_true, on_false):
    return _wrap(_np.where(pred, on_true, on_false))


def dynamic_slice(operand, start_indices, slice_sizes):
    a = _np.asarray(operand)
    sl = []
    for s, n, dim in zip(start_indices, slice_sizes, a.shape):
        s = min(max(int(s), 0), dim - n)              # XLA clamps the start
        sl.append(slice(s, s + n))
    return _wrap(a[tuple(sl)])
