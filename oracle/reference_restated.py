"""CPU restatement (NumPy FP64) of metMHN's training hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.  The
product (metmhn_b200) never does and has no CPU fallback.

It follows the reference ALGORITHM, not just its results: the same
reshape-C / flatten-F Kronecker shuffles, the same per-event operator terms,
the same (k+1)-sweep Jacobi resolvent solves and the same adjoint gradient
assembly.  It is written table-driven (one generic shuffle engine + the factor
tables of SURVEY.md Appendix A) instead of one function per factor.

Pinned against: tests/golden/*.npz, which tests/golden/make_golden.py produced
by executing the UNMODIFIED reference sources from /root/reference on a NumPy
stand-in for the JAX API (oracle/jax_shim) -- real JAX is not installable in
this image.  Independent second check: oracle/dense.py (explicit matrices,
numpy.linalg.solve, complex-step derivatives).

Reference files restated (all under /root/reference/metmhn):
  jx/kronvec.py, jx/vanilla.py, jx/likelihood.py, jx/one_event.py,
  regularized_optimization.py:11-298.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# generic shuffle engine  (kronvec.py:31-211: every k* factor is one of these two forms)
# --------------------------------------------------------------------------------------


def _mul(p, mult):
    """Per-pattern multiplier factor: reshape(-1,w) C-order, scale columns, flatten F-order."""
    w = len(mult)
    if w == 1:
        return p * mult[0]
    return (p.reshape((-1, w), order="C") * np.asarray(mult, dtype=float)).flatten(order="F")


def _shuffle(p, w):
    if w == 1:
        return p
    return p.reshape((-1, w), order="C").flatten(order="F")


def _trans(p, w, pairs, theta, diag, transpose):
    """Transition factor (k2ntt / k4ns / k4np / k4nm, kronvec.py:82-147): rate theta from
    pattern src to pattern dst for (src, dst) in pairs, -theta on the source diagonal if diag."""
    t = np.zeros((w, w))
    for s, d in pairs:
        t[s, d] += theta
        if diag:
            t[s, s] -= theta
    if transpose:
        t = t.T
    return (p.reshape((-1, w), order="C") @ t).flatten(order="F")


def diagnosis_theta(log_theta, log_d):
    """kronvec.py:7-21: subtract log_d[j] from every off-diagonal entry of column j."""
    out = log_theta - np.asarray(log_d)[None, :]
    idx = np.arange(log_theta.shape[0])
    out[idx, idx] = np.diagonal(log_theta)
    return out


def _cls(state, j):
    return int(state[2 * j]) + 2 * int(state[2 * j + 1])


_W = (1, 2, 2, 4)                    # vector width consumed per bit class (none, PT, MT, both)

# Appendix A.1: multipliers of event j != i, per term and class  (kronvec.py:223-228, 299-304, 370-376)
_OTHER = {
    "sync": lambda th: ([1.0], [1.0, 0.0], [1.0, 0.0], [1.0, 0.0, 0.0, th]),
    "prim": lambda th: ([1.0], [1.0, th], [1.0, 1.0], [1.0, th, 1.0, th]),
    "met":  lambda th: ([1.0], [1.0, 1.0], [1.0, th], [1.0, 1.0, th, th]),
}
# Appendix A.2: own factor; ("m", multipliers) or ("t", transition pairs)  (kronvec.py:236-239, 313-316, 384-388)
_OWN = {
    "sync": lambda th: (("m", [-th]), ("m", [-th, 0.0]), ("m", [-th, 0.0]), ("t", [(0, 3)])),
    "prim": lambda th: (("m", [-th]), ("t", [(0, 1)]), ("m", [-th, -th]), ("t", [(0, 1), (2, 3)])),
    "met":  lambda th: (("m", [-th]), ("m", [-th, -th]), ("t", [(0, 1)]), ("t", [(0, 2), (1, 3)])),
}
# pure-diagonal own factors used by kron_diag (kronvec.py:747-751, 795-798, 871-875)
_OWN_DIAG = {
    "sync": lambda th: ([-th], [-th, 0.0], [-th, 0.0], [-th, 0.0, 0.0, 0.0]),
    "prim": lambda th: ([-th], [-th, 0.0], [-th, -th], [-th, 0.0, -th, 0.0]),
    "met":  lambda th: ([-th], [-th, -th], [-th, 0.0], [-th, -th, 0.0, 0.0]),
}


def _joint_term(kind, log_theta, p, i, state, diag=True, transpose=False):
    """_kronvec_sync / _kronvec_prim / _kronvec_met (kronvec.py:214-397) incl. the
    short-circuits of the public wrappers (kronvec.py:283-287, 353-359, 425-431)."""
    n = log_theta.shape[0] - 1
    seeded = int(state[-1]) == 1
    ci = _cls(state, i)
    if kind == "sync":
        if (not diag) and ci != 3:
            return p * 0.0
    elif kind == "prim":
        if ((not diag) and int(state[2 * i]) == 0) or not seeded:
            return p * 0.0
    else:
        if ((not diag) and int(state[2 * i + 1]) == 0) or not seeded:
            return p * 0.0
    th = np.exp(log_theta[i, :])
    for j in range(n):
        c = _cls(state, j)
        if j != i:
            p = _mul(p, _OTHER[kind](th[j])[c])
        else:
            form, arg = _OWN[kind](th[i])[c]
            p = _mul(p, arg) if form == "m" else _trans(p, _W[c], arg, th[i], diag, transpose)
    if kind == "sync":
        if seeded:
            p = _mul(p, [1.0, 0.0])
    elif kind == "prim":
        p = _mul(p, [0.0, 1.0])
    else:
        p = _mul(p, [0.0, th[n]])
    return p


def kronvec_sync(log_theta, p, i, state, diag=True, transpose=False):
    return _joint_term("sync", log_theta, p, i, state, diag, transpose)


def kronvec_prim(log_theta, p, i, state, diag=True, transpose=False):
    return _joint_term("prim", log_theta, p, i, state, diag, transpose)


def kronvec_met(log_theta, p, i, state, diag=True, transpose=False):
    return _joint_term("met", log_theta, p, i, state, diag, transpose)


def kronvec_seed(log_theta, p, state, diag=True, transpose=False):
    """kronvec.py:434-496."""
    n = log_theta.shape[0] - 1
    seeded = int(state[-1]) == 1
    if (not diag) and not seeded:
        return p * 0.0
    th = np.exp(log_theta[n, :])
    for j in range(n):
        p = _mul(p, _OTHER["sync"](th[j])[_cls(state, j)])
    if seeded:
        return _trans(p, 2, [(0, 1)], th[n], diag, transpose)
    return -th[n] * p


def kronvec(log_theta, p, state, diag=True, transpose=False):
    """kronvec.py:499-539: Q p = sum_i (sync_i + prim_i + met_i) p + seed p."""
    n = log_theta.shape[0] - 1
    y = np.zeros_like(p)
    for i in range(n):
        y = y + kronvec_sync(log_theta, p, i, state, diag, transpose)
        y = y + kronvec_prim(log_theta, p, i, state, diag, transpose)
        y = y + kronvec_met(log_theta, p, i, state, diag, transpose)
    return y + kronvec_seed(log_theta, p, state, diag, transpose)


def kron_diag(log_theta, state, n_state):
    """kronvec.py:713-999: diagonal of the restricted joint Q."""
    n = log_theta.shape[0] - 1
    seeded = int(state[-1]) == 1
    y = np.zeros(2 ** n_state)
    for i in range(n):
        th = np.exp(log_theta[i, :])
        for kind in ("sync", "prim", "met"):
            if kind != "sync" and not seeded:
                continue
            d = np.ones(2 ** n_state)
            for j in range(n):
                c = _cls(state, j)
                d = _mul(d, (_OTHER[kind](th[j]) if j != i else _OWN_DIAG[kind](th[i]))[c])
            if kind == "sync":
                if seeded:
                    d = _mul(d, [1.0, 0.0])
            elif kind == "prim":
                d = _mul(d, [0.0, 1.0])
            else:
                d = _mul(d, [0.0, th[n]])
            y = y + d
    th = np.exp(log_theta[n, :])
    d = np.ones(2 ** n_state)
    for j in range(n):
        d = _mul(d, _OTHER["sync"](th[j])[_cls(state, j)])
    d = _mul(d, [-th[n], 0.0]) if seeded else -th[n] * d
    return y + d


def _diag_scal(side, log_d, state, p, own=None):
    """diag_scal_p / diag_scal_m (kronvec.py:574-602, 646-671) and, with own=i, the
    _partial_* variants whose own factor keeps only the patterns with event i set
    (kronvec.py:605-624, 674-701)."""
    n = log_d.shape[0] - 1
    d = np.exp(log_d)
    for j in range(n):
        c = _cls(state, j)
        if side == "p":
            mult = ([1.0], [1.0, d[j]], [1.0, 1.0], [1.0, d[j], 1.0, d[j]])[c]
            if own == j:
                mult = [0.0, d[j]] if c != 3 else [0.0, d[j], 0.0, d[j]]
        else:
            mult = ([1.0], [1.0, 1.0], [1.0, d[j]], [1.0, 1.0, d[j], d[j]])[c]
            if own == j:
                mult = [0.0, d[j]] if c != 3 else [0.0, 0.0, d[j], d[j]]
        p = _mul(p, mult)
    return _mul(p, [1.0, d[n]] if side == "p" else [0.0, d[n]])


def diag_scal_p(log_d_p, state, p):
    return _diag_scal("p", log_d_p, state, p)


def diag_scal_m(log_d_m, state, p):
    return _diag_scal("m", log_d_m, state, p)


def partial_diag_scal_p(log_d_p, state, p, i):
    """kronvec.py:630-644."""
    n = log_d_p.shape[0] - 1
    sel = int(state[2 * i]) + int(i == n)
    if sel == 0:
        return p * 0.0
    if sel == 1:
        return _diag_scal("p", log_d_p, state, p, own=i)
    out = diag_scal_p(log_d_p, state, p).reshape((-1, 2), order="F").copy()
    out[:, 0] = 0.0
    return out.ravel(order="F")


def partial_diag_scal_m(log_d_m, state, p, i):
    """kronvec.py:704-710."""
    n = log_d_m.shape[0] - 1
    sel = int(state[min(2 * n, 2 * i + 1)]) + int(i == n)
    if sel == 0:
        return p * 0.0
    if sel == 1:
        return _diag_scal("m", log_d_m, state, p, own=i)
    return diag_scal_m(log_d_m, state, p)


def obs_states(n_joint, state, pt_first=True):
    """kronvec.py:1056-1095: mask of joint states compatible with the first observation."""
    n = (len(state) - 1) // 2
    p = np.ones(2 ** n_joint)
    for i in range(n):
        c = _cls(state, i)
        if c == 0:
            continue
        if pt_first:
            mult = ([0.0, 1.0], [1.0, 1.0], [0.0, 1.0, 0.0, 1.0])[c - 1]
        else:
            mult = ([1.0, 1.0], [0.0, 1.0], [0.0, 0.0, 1.0, 1.0])[c - 1]
        p = _mul(p, mult)
    if int(state[-1]) == 1:
        p = _mul(p, [0.0, 1.0])
    return p


# --------------------------------------------------------------------------------------
# single-tumour MHN  (vanilla.py)
# --------------------------------------------------------------------------------------


def v_kronvec_i(log_theta, p, i, state, diag=True, transpose=False):
    """vanilla.py:21-75."""
    if (not diag) and int(state[i]) != 1:
        return 0.0 * p
    n = log_theta.shape[0]
    th = np.exp(log_theta[i, :])
    for j in range(n):
        if j != i:
            if int(state[j]) != 0:
                p = _mul(p, [1.0, th[j]])
        elif int(state[i]) == 0:
            p = -th[i] * p
        else:
            p = _trans(p, 2, [(0, 1)], th[i], diag, transpose)
    return p


def v_kronvec(log_theta, p, state, diag=True, transpose=False):
    """vanilla.py:78-106."""
    n = log_theta.shape[0]
    return np.sum([v_kronvec_i(log_theta, p, i, state, diag, transpose) for i in range(n)], axis=0)


def v_kron_diag(log_theta, state, diag):
    """vanilla.py:206-260."""
    n = log_theta.shape[0]
    tot = np.zeros_like(diag)
    for i in range(n):
        th = np.exp(log_theta[i, :])
        d = diag
        for j in range(n):
            if j != i:
                if int(state[j]) != 0:
                    d = _mul(d, [1.0, th[j]])
            elif int(state[i]) == 0:
                d = -th[i] * d
            else:
                d = _mul(d, [-th[i], 0.0])
        tot = tot + d
    return tot


def v_R_inv_vec(log_theta, x, state, d_rates=1.0, transpose=False):
    """vanilla.py:269-305: (d_rates - Q)^{-1} x by log2(len)+1 Jacobi sweeps."""
    k = int(np.log2(x.shape[0]))
    lidg = -1.0 / (v_kron_diag(log_theta, state, np.ones_like(x)) - d_rates)
    y = lidg * x
    for _ in range(k + 1):
        y = lidg * (v_kronvec(log_theta, y, state, False, transpose) + x)
    return y


def scal_d_pt(log_d_p, log_d_m, state, vec, own=None):
    """vanilla.py:125-167 (scal_d_pt, and _d_scal_d_pt when own is given)."""
    n = log_d_m.shape[0] - 1
    dp, dm = np.exp(log_d_p), np.exp(log_d_m)
    a, b = vec, vec
    for j in range(n):
        if own == j:
            a, b = _mul(a, [0.0, dp[j]]), _mul(b, [0.0, dm[j]])
        elif int(state[j]) == 1:
            a, b = _mul(a, [1.0, dp[j]]), _mul(b, [1.0, dm[j]])
    return _mul(a, [1.0, 0.0]), _mul(b, [0.0, dm[n]])


def d_scal_d_pt(log_d_p, log_d_m, state, vec, i):
    """vanilla.py:180-187."""
    n = log_d_p.shape[0] - 1
    sel = int(state[i]) + int(i == n)
    if sel == 0:
        return 0.0 * vec, 0.0 * vec
    if sel == 1:
        return scal_d_pt(log_d_p, log_d_m, state, vec, own=i)
    return 0.0 * vec, scal_d_pt(log_d_p, log_d_m, state, vec)[1]


def v_x_partial_D_y(log_d_p, log_d_m, state, x, y):
    """vanilla.py:190-203."""
    n = log_d_p.shape[0]
    res = np.zeros((n, 2))
    for i in range(n):
        a, b = d_scal_d_pt(log_d_p, log_d_m, state, y, i)
        res[i] = (np.dot(x, a), np.dot(x, b))
    return res[:, 0], res[:, 1]


def v_x_partial_Q_y(log_theta, x, y, state):
    """vanilla.py:328-393."""
    n = log_theta.shape[0]
    val = np.zeros((n, n))
    for i in range(n):
        z = x * v_kronvec_i(log_theta, y, i, state)
        for j in range(n):
            if j == i:
                val[i, i] = z.sum()
                if int(state[i]) != 0:
                    z = _shuffle(z, 2)
            elif int(state[j]) != 0:
                val[i, j] = z.reshape((-1, 2), order="C")[:, 1].sum()
                z = _shuffle(z, 2)
    d_diag = -val.sum(axis=0) + np.diagonal(val)
    return val, d_diag


def v_gradient(log_theta, state, p_0):
    """vanilla.py:396-418."""
    p_theta = v_R_inv_vec(log_theta, p_0, state)
    x = np.zeros_like(p_theta)
    x[-1] = 1.0 / p_theta[-1]
    x = v_R_inv_vec(log_theta, x, state, transpose=True)
    d_th, d_diag = v_x_partial_Q_y(log_theta, x, p_theta, state)
    return d_th, d_diag, p_theta


# --------------------------------------------------------------------------------------
# joint likelihood / gradient  (likelihood.py)
# --------------------------------------------------------------------------------------


def R_i_inv_vec(log_theta, log_d_p, log_d_m, x, state, state_size, transpose=False):
    """likelihood.py:231-262: (D_P + D_M - Q)^{-1} x by state_size+1 Jacobi sweeps."""
    one = np.ones_like(x)
    lidg = -1.0 / (kron_diag(log_theta, state, state_size)
                   - (diag_scal_p(log_d_p, state, one) + diag_scal_m(log_d_m, state, one)))
    y = lidg * x
    for _ in range(state_size + 1):
        y = lidg * (kronvec(log_theta, y, state, False, transpose) + x)
    return y


def _reduce_step(zs, zp, zm, c, own):
    """likelihood.py:25-107 (f0..f3 for j != i, t1/t12/t3 for j == i)."""
    w = _W[c]
    if own:
        if c in (1, 2):
            val = zs.reshape((-1, 2), order="C")[:, 0].sum() + zp.sum() + zm.sum()
        else:
            val = zs.sum() + zp.sum() + zm.sum()
    elif c == 0:
        val = 0.0
    elif c == 1:
        val = zp.reshape((-1, 2), order="C")[:, 1].sum()
    elif c == 2:
        val = zm.reshape((-1, 2), order="C")[:, 1].sum()
    else:
        val = (zs.reshape((-1, 4), order="C")[:, 3].sum()
               + zp.reshape((-1, 4), order="C")[:, [1, 3]].sum()
               + zm.reshape((-1, 4), order="C")[:, [2, 3]].sum())
    return _shuffle(zs, w), _shuffle(zp, w), _shuffle(zm, w), val


def x_partial_Q_y(log_theta, x, y, state):
    """likelihood.py:125-201."""
    n = log_theta.shape[0] - 1
    z = np.zeros((n + 1, n + 1))
    for i in range(n):
        zs = x * kronvec_sync(log_theta, y, i, state)
        zp = x * kronvec_prim(log_theta, y, i, state)
        zm = x * kronvec_met(log_theta, y, i, state)
        z[i, n] = zm.sum()
        for j in range(n):
            zs, zp, zm, z[i, j] = _reduce_step(zs, zp, zm, _cls(state, j), own=(j == i))
    zseed = x * kronvec_seed(log_theta, y, state)
    z[n, n] = zseed.sum()
    for j in range(n):
        c = _cls(state, j)
        z[n, j] = zseed.reshape((-1, 4), order="C")[:, 3].sum() if c == 3 else 0.0
        zseed = _shuffle(zseed, _W[c])
    return z


def x_partial_D_y(log_d_m, log_d_p, state, x, y):
    """likelihood.py:204-228 (argument order (log_d_m, log_d_p, ...) as in the reference)."""
    n = log_d_m.shape[0]
    d_dp, d_dm = np.zeros(n), np.zeros(n)
    for i in range(n):
        d_dp[i] = np.dot(x, partial_diag_scal_p(log_d_p, state, y, i))
        d_dm[i] = np.dot(x, partial_diag_scal_m(log_d_m, state, y, i))
    return d_dp, d_dm


def _cond_obs(vec_scaled, state_joint, n_joint, n_single, pt_first):
    """likelihood.py:265-283 / 559-562 / 599-602."""
    mask = obs_states(n_joint, state_joint, pt_first)
    size = 2 ** (n_single - 1)
    inds = np.where(mask == 1.0)[0][:size]
    return np.append(np.zeros(size), vec_scaled[inds]), mask, inds


def _theta_pt(log_theta, log_d_p):
    t = log_theta.copy()
    t[:-1, -1] = 0.0
    return diagnosis_theta(t, log_d_p)


def _e0(k):
    p = np.zeros(2 ** k)
    p[0] = 1.0
    return p


def _lp_coupled(order, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    """likelihood.py:286-384 (order 0 / 1 / anything else)."""
    n_joint = n_prim + n_met - 1
    y = R_i_inv_vec(log_theta, log_d_p, log_d_m, _e0(n_joint), state_joint, n_joint)
    tot = 0.0
    if order in (0, 1):
        v, _, _ = _cond_obs(diag_scal_p(log_d_p, state_joint, y), state_joint, n_joint, n_met, True)
        met = np.append(state_joint[1::2], 1)
        tot += v_R_inv_vec(diagnosis_theta(log_theta, log_d_m), v, met)[-1]
    if order != 1:
        v, _, _ = _cond_obs(diag_scal_m(log_d_m, state_joint, y), state_joint, n_joint, n_prim, False)
        tot += v_R_inv_vec(_theta_pt(log_theta, log_d_p), v, state_joint[0::2])[-1]
    return np.log(tot)


def _lp_coupled_0(log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    return _lp_coupled(0, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met)


def _lp_coupled_1(log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    return _lp_coupled(1, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met)


def _lp_coupled_2(log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    return _lp_coupled(2, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met)


def _lp_prim_obs(log_theta, log_d_p, state_pt, n_prim):
    """likelihood.py:387-405."""
    p = v_R_inv_vec(_theta_pt(log_theta, log_d_p), _e0(n_prim), state_pt, np.ones(2 ** n_prim))
    return np.log(p[-1])


def _lp_prim_obs_az(log_theta):
    """likelihood.py:408-416."""
    return np.log(1.0 / (1.0 + np.sum(np.exp(np.diag(log_theta)))))


def _lp_met_obs(log_theta, log_d_p, log_d_m, state_mt, n_met):
    """likelihood.py:419-438."""
    a, b = scal_d_pt(log_d_p, log_d_m, state_mt, np.ones(2 ** n_met))
    d_rates = a + b
    p = v_R_inv_vec(log_theta, _e0(n_met), state_mt, d_rates, False)
    return np.log(p[-1] * d_rates[-1])


def _grad_prim_obs(log_theta, log_d_p, state_prim, n_prim):
    """likelihood.py:441-461."""
    d_th, d_dp, p = v_gradient(_theta_pt(log_theta, log_d_p), state_prim, _e0(n_prim))
    d_th[:-1, -1] = 0.0
    return np.log(p[-1]), d_th, d_dp


def _grad_prim_obs_az(log_theta):
    """likelihood.py:464-478."""
    th = np.exp(np.diag(log_theta))
    log_p = np.log(1.0 / (1.0 + th.sum()))
    d_th = 1.0 / np.exp(log_p) * np.diag(-th / (1.0 + th.sum()) ** 2)
    return log_p, d_th, np.zeros(log_theta.shape[0])


def _grad_met_obs(log_theta, log_d_p, log_d_m, state_met, n_met):
    """likelihood.py:481-512."""
    a, b = scal_d_pt(log_d_p, log_d_m, state_met, np.ones(2 ** n_met))
    d_rates = a + b
    p = v_R_inv_vec(log_theta, _e0(n_met), state_met, d_rates, False)
    score = p[-1]
    q = np.zeros(2 ** n_met)
    q[-1] = 1.0 / score
    _, d_dm_1 = v_x_partial_D_y(log_d_p, log_d_m, state_met, q / d_rates[-1], p)
    q = v_R_inv_vec(log_theta, q, state_met, d_rates, True)
    d_dp, d_dm_2 = v_x_partial_D_y(log_d_p, log_d_m, state_met, q, p)
    d_th, _ = v_x_partial_Q_y(log_theta, q, p, state_met)
    return np.log(score * d_rates[-1]), d_th, -d_dp, d_dm_1 - d_dm_2


def _q_inv_deriv_pth(log_theta, log_d_p, log_d_m, q, p, state_joint, n_joint):
    """likelihood.py:516-537."""
    q = R_i_inv_vec(log_theta, log_d_p, log_d_m, q, state_joint, n_joint, transpose=True)
    g_2 = x_partial_Q_y(log_theta, q, p, state_joint)
    d_dp_2, d_dm_2 = x_partial_D_y(log_d_m, log_d_p, state_joint, q, p)
    return g_2, d_dp_2, d_dm_2


def _marginal(pt_first, log_theta, log_d_p, log_d_m, y, state_joint, n_joint, n_single):
    """marginal_obs_pt_first / marginal_obs_mt_first (likelihood.py:540-620)."""
    if pt_first:
        scaled = diag_scal_p(log_d_p, state_joint, y)
        th2 = diagnosis_theta(log_theta, log_d_m)
        single = np.append(state_joint[1::2], 1)
    else:
        scaled = diag_scal_m(log_d_m, state_joint, y)
        th2 = _theta_pt(log_theta, log_d_p)
        single = state_joint[0::2]
    v, mask, inds = _cond_obs(scaled, state_joint, n_joint, n_single, pt_first)
    g_1, d_diag, p2 = v_gradient(th2, single, v)
    if not pt_first:
        g_1[:-1, -1] = 0.0
    exp_score = p2[-1]
    q = np.zeros(2 ** n_single)
    q[-1] = 1.0 / exp_score
    q = v_R_inv_vec(th2, q, single, transpose=True)
    p = mask * 0.0
    p[inds] = q[2 ** (n_single - 1):]
    d_dp, d_dm = x_partial_D_y(log_d_m, log_d_p, state_joint, p, y)
    if pt_first:
        return exp_score, g_1, d_dp, d_diag, p          # (score, g_1, d_dp_1, d_dm_1, p)
    return exp_score, g_1, d_diag, d_dm, p


def _g_coupled(order, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    """_g_coupled_0 / _1 / _2 (likelihood.py:623-730)."""
    n_joint = n_prim + n_met - 1
    y = R_i_inv_vec(log_theta, log_d_p, log_d_m, _e0(n_joint), state_joint, n_joint)
    if order == 1:
        s, g_1, dp_1, dm_1, p = _marginal(True, log_theta, log_d_p, log_d_m, y, state_joint, n_joint, n_met)
        p = diag_scal_p(log_d_p, state_joint, p)
    elif order != 0:
        s, g_1, dp_1, dm_1, p = _marginal(False, log_theta, log_d_p, log_d_m, y, state_joint, n_joint, n_prim)
        p = diag_scal_m(log_d_m, state_joint, p)
    else:
        s_pf, g_pf, dp_pf, dm_pf, p_pf = _marginal(True, log_theta, log_d_p, log_d_m, y, state_joint, n_joint, n_met)
        s_mf, g_mf, dp_mf, dm_mf, p_mf = _marginal(False, log_theta, log_d_p, log_d_m, y, state_joint, n_joint, n_prim)
        s = s_pf + s_mf
        p = (diag_scal_p(log_d_p, state_joint, p_pf) * s_pf / s
             + diag_scal_m(log_d_m, state_joint, p_mf) * s_mf / s)
        g_1 = (g_pf * s_pf + g_mf * s_mf) / s
        dp_1 = (dp_pf * s_pf + dp_mf * s_mf) / s
        dm_1 = (dm_pf * s_pf + dm_mf * s_mf) / s
    g_2, dp_2, dm_2 = _q_inv_deriv_pth(log_theta, log_d_p, log_d_m, p, y, state_joint, n_joint)
    return np.log(s), g_1 + g_2, dp_1 - dp_2, dm_1 - dm_2


def _g_coupled_0(log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    return _g_coupled(0, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met)


def _g_coupled_1(log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    return _g_coupled(1, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met)


def _g_coupled_2(log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met):
    return _g_coupled(2, log_theta, log_d_p, log_d_m, state_joint, n_prim, n_met)


# --------------------------------------------------------------------------------------
# paired patient with only the seeding bit set  (one_event.py)
# --------------------------------------------------------------------------------------


def _one_small_Q(log_theta):
    """one_event.py:10-25."""
    base = np.diagonal(log_theta)
    b_r = np.exp(base[:-1])
    e_seed = np.exp(log_theta[:-1, -1]) + 1.0
    return np.array([[-np.exp(base).sum(), 0.0],
                     [np.exp(log_theta[-1, -1]), -np.sum(b_r * e_seed)]])


def _one_solve(log_theta, x, dp_le, dm_le, transpose=False):
    """one_event.py:56-85."""
    R = np.diag([1.0, dp_le + dm_le]) - _one_small_Q(log_theta)
    b = np.array(x, dtype=float)
    if not transpose:
        b[0] /= R[0, 0]
        b[1] -= b[0] * R[1, 0]
        b[1] /= R[1, 1]
    else:
        b[1] /= R[1, 1]
        b[0] -= b[1] * R[1, 0]
        b[0] /= R[0, 0]
    return b


def _one_x_partial_Q_y(log_theta, x, y):
    """one_event.py:88-113."""
    n = log_theta.shape[0]
    z = np.zeros_like(log_theta)
    for i in range(n):
        t_ii, t_iM = np.exp(log_theta[i, i]), np.exp(log_theta[i, -1])
        z[i, i] = -t_ii * (x[0] * y[0] + (1.0 + t_iM) * x[1] * y[1])
        z[i, -1] = -t_ii * t_iM * x[1] * y[1]
    t_MM = np.exp(log_theta[-1, -1])
    z[-1, -1] = x @ np.array([[-t_MM, 0.0], [t_MM, 0.0]]) @ y
    return z


def _one_marginal(pt_first, log_theta, log_d_p, log_d_m, y, state_joint):
    """one_event.py:229-303."""
    n = log_theta.shape[0]
    if pt_first:
        le = np.exp(log_d_p[-1])
        th2 = diagnosis_theta(log_theta, log_d_m)
        single = np.append(state_joint[1::2], 1)
    else:
        le = np.exp(log_d_m[-1])
        th2 = _theta_pt(log_theta, log_d_p)
        single = state_joint[0::2]
    v = np.array([0.0, y[-1] * le])
    g_1, d_diag, p2 = v_gradient(th2, single, v)
    if not pt_first:
        g_1[:-1, -1] = 0.0
    s = p2[-1]
    q = v_R_inv_vec(th2, np.array([0.0, 1.0 / s]), single, transpose=True)
    p = q * np.array([0.0, le])
    d_own = np.zeros(n)
    d_own[-1] = np.dot(p, y)
    if pt_first:
        return s, g_1, d_own, d_diag, p
    return s, g_1, d_diag, d_own, p


def _one_g_coupled(order, log_theta, log_d_p, log_d_m, state_joint):
    """one_event.py:307-408."""
    n = log_theta.shape[0]
    dp_le, dm_le = np.exp(log_d_p[-1]), np.exp(log_d_m[-1])
    y = _one_solve(log_theta, np.array([1.0, 0.0]), dp_le, dm_le)
    if order == 1:
        s, g_1, dp_1, dm_1, p = _one_marginal(True, log_theta, log_d_p, log_d_m, y, state_joint)
    elif order != 0:
        s, g_1, dp_1, dm_1, p = _one_marginal(False, log_theta, log_d_p, log_d_m, y, state_joint)
    else:
        s_pf, g_pf, dp_pf, dm_pf, p_pf = _one_marginal(True, log_theta, log_d_p, log_d_m, y, state_joint)
        s_mf, g_mf, dp_mf, dm_mf, p_mf = _one_marginal(False, log_theta, log_d_p, log_d_m, y, state_joint)
        s = s_pf + s_mf
        p = (p_pf * s_pf + p_mf * s_mf) / s
        g_1 = (g_pf * s_pf + g_mf * s_mf) / s
        dp_1 = (dp_pf * s_pf + dp_mf * s_mf) / s
        dm_1 = (dm_pf * s_pf + dm_mf * s_mf) / s
    q = _one_solve(log_theta, p, dp_le, dm_le, transpose=True)
    g_2 = _one_x_partial_Q_y(log_theta, q, y)
    dp_2, dm_2 = np.zeros(n), np.zeros(n)
    dm_2[-1] = np.dot(q * np.array([0.0, dm_le]), y)
    dp_2[-1] = np.dot(q * np.array([0.0, dp_le]), y)
    return np.log(s), g_1 + g_2, dp_1 - dp_2, dm_1 - dm_2


def _one_lp_coupled(order, log_theta, log_d_p, log_d_m, state_joint):
    """one_event.py:141-226."""
    dp_le, dm_le = np.exp(log_d_p[-1]), np.exp(log_d_m[-1])
    y = _one_solve(log_theta, np.array([1.0, 0.0]), dp_le, dm_le)
    tot = 0.0
    if order in (0, 1):
        met = np.append(state_joint[1::2], 1)
        tot += v_R_inv_vec(diagnosis_theta(log_theta, log_d_m), np.array([0.0, y[-1] * dp_le]), met)[-1]
    if order != 1:
        tot += v_R_inv_vec(_theta_pt(log_theta, log_d_p), np.array([0.0, y[-1] * dm_le]), state_joint[0::2])[-1]
    return np.log(tot)


# --------------------------------------------------------------------------------------
# dataset level  (regularized_optimization.py)
# --------------------------------------------------------------------------------------


def patient_value_grad(log_theta, log_d_p, log_d_m, row, want_grad=True):
    """Dispatch of one data row exactly as regularized_optimization.py:75-119 / 187-254.
    Returns (kind_is_type0, logp, d_th, d_dp, d_dm); grads are None if want_grad is False;
    rows with an unknown type return None."""
    n_tot = log_theta.shape[0]
    n_mut = n_tot - 1
    z = np.zeros(n_tot)
    typ = int(row[-1])
    if typ in (0, 1):
        st = np.asarray(row[0:2 * n_tot - 1:2])
        k = int(st.sum())
        if typ == 0 and k == 0:
            if want_grad:
                lp, g, dp = _grad_prim_obs_az(log_theta)
                return typ == 0, lp, g, dp, z
            return True, _lp_prim_obs_az(log_theta), None, None, None
        if want_grad:
            lp, g, dp = _grad_prim_obs(log_theta, log_d_p, st, k)
            return typ == 0, lp, g, dp, z
        return typ == 0, _lp_prim_obs(log_theta, log_d_p, st, k), None, None, None
    if typ == 2:
        st = np.append(np.asarray(row[1:2 * n_tot - 1:2]), 1)
        k = int(st.sum())
        if want_grad:
            lp, g, dp, dm = _grad_met_obs(log_theta, log_d_p, log_d_m, st, k)
            return False, lp, g, dp, dm
        return False, _lp_met_obs(log_theta, log_d_p, log_d_m, st, k), None, None, None
    if typ == 3:
        st = np.asarray(row[0:2 * n_mut + 1])
        n_prim = int(st[::2].sum())
        n_met = int(st[1::2].sum() + 1)
        order = int(row[-2])
        order = order if order in (0, 1) else 2
        if n_prim + n_met - 1 == 1:
            if want_grad:
                return (False,) + _one_g_coupled(order, log_theta, log_d_p, log_d_m, st)
            return False, _one_lp_coupled(order, log_theta, log_d_p, log_d_m, st), None, None, None
        if want_grad:
            return (False,) + _g_coupled(order, log_theta, log_d_p, log_d_m, st, n_prim, n_met)
        return False, _lp_coupled(order, log_theta, log_d_p, log_d_m, st, n_prim, n_met), None, None, None
    return None


def _weights(dat, perc_met):
    """regularized_optimization.py:121-128 / 256-262."""
    n_em = float(np.sum(dat[:, -3].astype(np.int64)))
    n_nm = dat.shape[0] - n_em
    w = perc_met * n_nm / ((1 - perc_met) * n_em) if n_em * n_nm != 0 else 1.0
    return w, w * n_em + n_nm


def score(log_theta, log_d_p, log_d_m, dat, perc_met):
    """regularized_optimization.py:55-130."""
    s_em, s_pt = 0.0, 0.0
    for i in range(dat.shape[0]):
        r = patient_value_grad(log_theta, log_d_p, log_d_m, dat[i], want_grad=False)
        if r is None:
            continue
        if r[0]:
            s_pt += r[1]
        else:
            s_em += r[1]
    w, n_full = _weights(dat, perc_met)
    return (w * s_em + s_pt) / n_full


def score_and_grad(log_theta, log_d_p, log_d_m, dat, perc_met):
    """regularized_optimization.py:163-267."""
    n_tot = log_theta.shape[0]
    s_em, s_pt = 0.0, 0.0
    g_em, g_pt = np.zeros((n_tot, n_tot)), np.zeros((n_tot, n_tot))
    dp_em, dp_pt, dm_em = np.zeros(n_tot), np.zeros(n_tot), np.zeros(n_tot)
    for i in range(dat.shape[0]):
        r = patient_value_grad(log_theta, log_d_p, log_d_m, dat[i], want_grad=True)
        if r is None:
            continue
        is0, lp, g, dp, dm = r
        if is0:
            s_pt += lp
            g_pt += g
            dp_pt += dp
        else:
            s_em += lp
            g_em += g
            dp_em += dp
            dm_em += dm
    w, n_full = _weights(dat, perc_met)
    return ((w * s_em + s_pt) / n_full, (w * g_em + g_pt) / n_full,
            (w * dp_em + dp_pt) / n_full, w * dm_em / n_full)


def L1(theta, eps=1e-5):
    """regularized_optimization.py:11-18."""
    t = np.array(theta, dtype=float)
    if t.ndim == 2:
        np.fill_diagonal(t, 0.0)
    return np.sum(np.sqrt(t ** 2 + eps))


def L1_(theta, eps=1e-5):
    """regularized_optimization.py:21-28."""
    t = np.array(theta, dtype=float)
    if t.ndim == 2:
        np.fill_diagonal(t, 0.0)
    return t.flatten() / np.sqrt(t.flatten() ** 2 + eps)


def sym_penal(log_theta, eps=1e-5):
    """regularized_optimization.py:31-35."""
    n = log_theta.shape[0]
    t = np.array(log_theta, dtype=float)
    np.fill_diagonal(t, 0.0)
    return 0.5 * (np.sum(np.sqrt(t ** 2 + t.T ** 2 - t * t.T + eps)) - n * np.sqrt(eps))


def sym_penal_(log_theta, eps=1e-5):
    """regularized_optimization.py:38-43."""
    t = np.array(log_theta, dtype=float)
    np.fill_diagonal(t, 0.0)
    return ((2 * t - t.T) / (2 * np.sqrt(t ** 2 + t.T ** 2 - t * t.T + eps))).flatten()


def _unpack(params, n_total):
    params = np.asarray(params, dtype=float)
    return (params[0:n_total ** 2].reshape((n_total, n_total)),
            params[n_total ** 2:n_total * (n_total + 1)], params[n_total * (n_total + 1):])


def symmetric_penal(params, n_total, eps=1e-5):
    """regularized_optimization.py:46-52."""
    th, dp, dm = _unpack(params, n_total)
    return (sym_penal(th) + L1(dp) + L1(dm),
            np.concatenate((sym_penal_(th), L1_(dp), L1_(dm))))


def score_reg(params, dat, perc_met, penal, w_penal):
    """regularized_optimization.py:133-160."""
    n_total = (dat.shape[1] - 3) // 2 + 1
    th, dp, dm = _unpack(params, n_total)
    pen, _ = penal(params, n_total)
    return -score(th, dp, dm, dat, perc_met) + w_penal * pen


def score_and_grad_reg(params, dat, perc_met, penal, w_penal):
    """regularized_optimization.py:270-298."""
    n_total = (dat.shape[1] - 3) // 2 + 1
    th, dp, dm = _unpack(params, n_total)
    s, g, gdp, gdm = score_and_grad(th, dp, dm, dat, perc_met)
    pen, pen_ = penal(params, n_total)
    return -s + w_penal * pen, -np.concatenate((g.flatten(), gdp, gdm)) + w_penal * pen_
