"""Per-patient entry points mirroring `metmhn/jx/likelihood.py` (and `one_event.py`), for tests and
notebooks that call them directly.  Each call builds a one-row dataset and runs the same CUDA path as
the dataset-level functions; no special case is needed for the seeding-only paired patient that the
reference routes to `one_event.py` (it is the k = 1 limit of the general lattice).
"""
from __future__ import annotations

import numpy as np

from ._lib import Handle


def _row(n, pt=None, mt=None, seeding=1, order=-99, typ=0):
    r = np.zeros(2 * n + 3, dtype=np.int8)
    if pt is not None:
        r[0:2 * n:2] = np.asarray(pt)[:n]
    if mt is not None:
        r[1:2 * n:2] = np.asarray(mt)[:n]
    r[2 * n], r[2 * n + 1], r[2 * n + 2] = seeding, order, typ
    return r[None, :]


def _run(log_theta, log_d_p, log_d_m, row, want_grad):
    n_tot = np.asarray(log_theta).shape[0]
    params = np.concatenate([np.asarray(log_theta, float).ravel(), np.asarray(log_d_p, float),
                             np.asarray(log_d_m, float)])
    h = Handle(row)
    try:
        s, g = h.eval_weighted(params, 1.0, 1.0, want_grad=want_grad)
    finally:
        h.close()
    if not want_grad:
        return s
    sq = n_tot * n_tot
    return s, g[:sq].reshape(n_tot, n_tot).copy(), g[sq:sq + n_tot].copy(), g[sq + n_tot:].copy()


def _coupled(order, want_grad, log_theta, log_d_p, log_d_m, state_joint):
    n = np.asarray(log_theta).shape[0] - 1
    st = np.asarray(state_joint)
    row = np.zeros((1, 2 * n + 3), dtype=np.int8)
    row[0, :2 * n + 1] = st[:2 * n + 1]
    row[0, -2], row[0, -1] = order, 3
    return _run(log_theta, log_d_p, log_d_m, row, want_grad)


def _g_coupled_0(log_theta, log_d_p, log_d_m, state_joint, n_prim=None, n_met=None):
    """likelihood.py:623 / one_event.py:307"""
    return _coupled(0, True, log_theta, log_d_p, log_d_m, state_joint)


def _g_coupled_1(log_theta, log_d_p, log_d_m, state_joint, n_prim=None, n_met=None):
    """likelihood.py:665 / one_event.py:346"""
    return _coupled(1, True, log_theta, log_d_p, log_d_m, state_joint)


def _g_coupled_2(log_theta, log_d_p, log_d_m, state_joint, n_prim=None, n_met=None):
    """likelihood.py:700 / one_event.py:379"""
    return _coupled(2, True, log_theta, log_d_p, log_d_m, state_joint)


def _lp_coupled_0(log_theta, log_d_p, log_d_m, state_joint, n_prim=None, n_met=None):
    """likelihood.py:286"""
    return _coupled(0, False, log_theta, log_d_p, log_d_m, state_joint)


def _lp_coupled_1(log_theta, log_d_p, log_d_m, state_joint, n_prim=None, n_met=None):
    """likelihood.py:320"""
    return _coupled(1, False, log_theta, log_d_p, log_d_m, state_joint)


def _lp_coupled_2(log_theta, log_d_p, log_d_m, state_joint, n_prim=None, n_met=None):
    """likelihood.py:353"""
    return _coupled(2, False, log_theta, log_d_p, log_d_m, state_joint)


def _grad_prim_obs(log_theta, log_d_p, state_prim, n_prim=None):
    """likelihood.py:442: returns (log p, d_theta, d_d_p)."""
    n = np.asarray(log_theta).shape[0] - 1
    st = np.asarray(state_prim)
    row = _row(n, pt=st[:n], seeding=int(st[n]), typ=1)
    s, g, dp, _ = _run(log_theta, log_d_p, np.zeros(n + 1), row, True)
    return s, g, dp


def _lp_prim_obs(log_theta, log_d_p, state_pt, n_prim=None):
    """likelihood.py:387"""
    n = np.asarray(log_theta).shape[0] - 1
    st = np.asarray(state_pt)
    return _run(log_theta, log_d_p, np.zeros(n + 1), _row(n, pt=st[:n], seeding=int(st[n]), typ=1), False)


def _grad_prim_obs_az(log_theta):
    """likelihood.py:465"""
    n = np.asarray(log_theta).shape[0] - 1
    s, g, dp, _ = _run(log_theta, np.zeros(n + 1), np.zeros(n + 1), _row(n, seeding=0, typ=0), True)
    return s, g, dp


def _lp_prim_obs_az(log_theta):
    """likelihood.py:408"""
    n = np.asarray(log_theta).shape[0] - 1
    return _run(log_theta, np.zeros(n + 1), np.zeros(n + 1), _row(n, seeding=0, typ=0), False)


def _grad_met_obs(log_theta, log_d_p, log_d_m, state_met, n_met=None):
    """likelihood.py:482: returns (log p, d_theta, d_d_p, d_d_m)."""
    n = np.asarray(log_theta).shape[0] - 1
    return _run(log_theta, log_d_p, log_d_m, _row(n, mt=np.asarray(state_met)[:n], typ=2), True)


def _lp_met_obs(log_theta, log_d_pt, log_d_mt, state_mt, n_met=None):
    """likelihood.py:419"""
    n = np.asarray(log_theta).shape[0] - 1
    return _run(log_theta, log_d_pt, log_d_mt, _row(n, mt=np.asarray(state_mt)[:n], typ=2), False)
