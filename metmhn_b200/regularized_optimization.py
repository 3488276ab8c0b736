"""Drop-in mirror of the reference's `metmhn/regularized_optimization.py` call surface.

Same names, argument meaning and packing as the reference (file:line cited per function); the
per-patient likelihood and adjoint gradient run on the GPU through the C-ABI of
include/metmhn_b200.h.  The dataset is preprocessed and uploaded once and cached, because SciPy's
L-BFGS-B calls `score_and_grad_reg` with the same `dat` at every iteration
(`regularized_optimization.py:328`).  There is no CPU fallback.
"""
from __future__ import annotations

import hashlib
from typing import Callable

import numpy as np

from ._lib import Handle

_CACHE: dict = {}
_CACHE_MAX = 8


def dataset_handle(dat, device: int = 0) -> Handle:
    """Handle for `dat` on `device`, cached by content (a 128-bit BLAKE2 digest of the int8 matrix: L-BFGS-B passes the same
    `dat` at every iteration; callers that evaluate many times should pass the Handle itself and skip the hashing)."""
    if isinstance(dat, Handle):
        return dat
    raw = np.asarray(dat)
    if raw.dtype != np.int8:
        if raw.size and (raw.min() < -128 or raw.max() > 127):
            raise ValueError("dat entries must fit int8 (genotypes 0/1, order, type)")
        if raw.dtype.kind == "f" and raw.size and not np.all(raw == np.round(raw)):
            raise ValueError("dat entries must be integers")
    a = np.ascontiguousarray(raw, dtype=np.int8)
    key = (a.shape, hashlib.blake2b(a.view(np.uint8).reshape(-1), digest_size=16).digest(), device)
    h = _CACHE.get(key)
    if h is None:
        if len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))      # only the cache's reference goes; a handle somebody still holds stays valid
        h = _CACHE[key] = Handle(a, device=device)
    return h


def clear_cache():
    _CACHE.clear()


def _pack(log_theta, log_d_p, log_d_m):
    return np.concatenate([np.asarray(log_theta, dtype=np.float64).ravel(),
                           np.asarray(log_d_p, dtype=np.float64).ravel(),
                           np.asarray(log_d_m, dtype=np.float64).ravel()])


def _unpack(vec, n_total):
    vec = np.asarray(vec, dtype=np.float64)
    sq = n_total * n_total
    return vec[:sq].reshape(n_total, n_total), vec[sq:sq + n_total], vec[sq + n_total:]


# ---- penalties (host side, O(n^2); regularized_optimization.py:11-52) ---------------------------------

def L1(theta, eps: float = 1e-05):
    """Smoothed L1 norm; the diagonal of a matrix argument is not penalised (:11-18)."""
    t = np.array(theta, dtype=np.float64)
    if t.ndim == 2:
        t[np.diag_indices(t.shape[0])] = 0.0
    return np.sqrt(t * t + eps).sum()


def L1_(theta, eps: float = 1e-05):
    """Derivative of L1, flattened (:21-28)."""
    t = np.array(theta, dtype=np.float64)
    if t.ndim == 2:
        t[np.diag_indices(t.shape[0])] = 0.0
    t = t.ravel()
    return t / np.sqrt(t * t + eps)


def sym_penal(log_theta, eps: float = 1e-05):
    """Symmetrised group penalty over the pairs (theta_ij, theta_ji) (:31-35)."""
    t = np.array(log_theta, dtype=np.float64)
    n = t.shape[0]
    t[np.diag_indices(n)] = 0.0
    pair = np.sqrt(t * t + t.T * t.T - t * t.T + eps)
    return 0.5 * (pair.sum() - n * np.sqrt(eps))


def sym_penal_(log_theta, eps: float = 1e-05):
    """Derivative of sym_penal, flattened (:38-43)."""
    t = np.array(log_theta, dtype=np.float64)
    t[np.diag_indices(t.shape[0])] = 0.0
    pair = np.sqrt(t * t + t.T * t.T - t * t.T + eps)
    return ((2.0 * t - t.T) / (2.0 * pair)).ravel()


def symmetric_penal(params, n_total: int, eps: float = 1e-05):
    """(penalty, gradient) for the packed parameter vector (:46-52)."""
    th, dp, dm = _unpack(params, n_total)
    value = sym_penal(th, eps) + L1(dp, eps) + L1(dm, eps)
    grad = np.concatenate([sym_penal_(th, eps), L1_(dp, eps), L1_(dm, eps)])
    return value, grad


# ---- likelihood (GPU) -----------------------------------------------------------------------------------

def score(log_theta, log_d_p, log_d_m, dat, perc_met: float):
    """Weighted mean log-likelihood of `dat` (reference :55-130)."""
    return dataset_handle(dat).value(_pack(log_theta, log_d_p, log_d_m), perc_met)


def score_and_grad(log_theta, log_d_p, log_d_m, dat, perc_met: float):
    """(score, d_theta (n+1)x(n+1), d_d_p, d_d_m) -- BASELINE's "value_grad" (reference :163-267)."""
    h = dataset_handle(dat)
    s, g = h.value_grad(_pack(log_theta, log_d_p, log_d_m), perc_met)
    gth, gdp, gdm = _unpack(g, h.n_tot)
    return s, gth.copy(), gdp.copy(), gdm.copy()


def score_reg(params, dat, perc_met: float, penal: Callable, w_penal: float):
    """Negative penalised log-likelihood (reference :133-160)."""
    h = dataset_handle(dat)
    pen, _ = penal(params, h.n_tot)
    return np.array(-h.value(params, perc_met) + w_penal * pen)


def score_and_grad_reg(params, dat, perc_met: float, penal: Callable, w_penal: float):
    """(f, g) handed to L-BFGS-B: -score + lambda*penalty and its gradient (reference :270-298)."""
    h = dataset_handle(dat)
    s, g = h.value_grad(params, perc_met)
    pen, pen_ = penal(params, h.n_tot)
    return np.array(-s + w_penal * pen), -g + w_penal * np.asarray(pen_)


LAST_FIT: dict = {}


def learn_mhn(th_init, dp_init, dm_init, dat, perc_met: float, penal: Callable, w_penal: float,
              opt_iter: int = 1e05, opt_ftol: float = 1e-04, opt_v: bool = True, optimizer: str | None = None):
    """Fit a metMHN (reference :301-334).

    optimizer="scipy": SciPy's `minimize(method="L-BFGS-B")` calls `score_and_grad_reg` exactly as the reference drives it.
    optimizer="native": the same algorithm (L-BFGS-B without bounds) inside the library, `mmh_learn`: likelihood, gradient
    AND penalty on the device, one C call per fit -- no Python or SciPy work per iteration (SciPy >= 1.15 spends several
    milliseconds per iteration in its own L-BFGS-B at these parameter counts, more than a LUAD evaluation takes on a B200).
    Needs the built-in `symmetric_penal`.  Default (None): "native" when `penal` is `symmetric_penal`, else "scipy".
    Both reach the same optimum; their iterates differ in the last digits, so the number of iterations can differ by a few.
    `LAST_FIT` records {optimizer, f, iterations, evaluations} of the last call."""
    h = dataset_handle(dat)
    n_total = np.asarray(th_init).shape[0]
    x0 = _pack(th_init, dp_init, dm_init)
    if optimizer is None:
        optimizer = "native" if penal is symmetric_penal else "scipy"
    if optimizer == "native":
        if penal is not symmetric_penal:
            raise ValueError("optimizer='native' evaluates symmetric_penal on the device; pass optimizer='scipy' for another penalty")
        x, f, it, ev = h.learn(x0, perc_met, w_penal, eps=1e-05, max_iter=int(opt_iter), ftol=opt_ftol)
        LAST_FIT.clear()
        LAST_FIT.update(optimizer="native", f=f, iterations=it, evaluations=ev)
        th, dp, dm = _unpack(x, n_total)
        return th.copy(), dp.copy(), dm.copy()
    if optimizer != "scipy":
        raise ValueError("optimizer must be 'native', 'scipy' or None")
    import scipy.optimize as opt

    options = {"maxiter": int(opt_iter), "ftol": opt_ftol}
    import scipy
    if tuple(int(v) for v in scipy.__version__.split(".")[:2]) < (1, 15):
        options["disp"] = opt_v           # SciPy >= 1.15 rewrote L-BFGS-B and no longer takes `disp`
    res = opt.minimize(fun=score_and_grad_reg, jac=True, x0=x0, method="L-BFGS-B",
                       args=(h, perc_met, penal, w_penal), options=options)
    LAST_FIT.clear()
    LAST_FIT.update(optimizer="scipy", f=float(res.fun), iterations=int(res.nit), evaluations=int(res.nfev))
    th, dp, dm = _unpack(res.x, n_total)
    return th.copy(), dp.copy(), dm.copy()
