"""Second CPU formulation of the hot path: the DIRECT subset-lattice form.

TEST INFRASTRUCTURE ONLY (same rules as reference_restated.py).

This is the executable specification of what the CUDA kernels compute
(DESIGN.md section 3): explicit transition rates on the subset lattice, an exact
triangular solve by popcount levels instead of Jacobi sweeps, a canonical bit
layout (PT bits low, MT bits high, shared events first), the joint space split
into its pre-seeding diagonal sub-lattice and the post-seeding Kronecker-sum
part, and the gradient assembled from marginal statistics.  It shares no code
with reference_restated.py, so agreement of the two (tests/test_oracle.py) checks
both the restatement and the derivation the kernels rely on.  It is vectorised
per level and is the checker used at sizes where the literal restatement
((k+1) k^2 shuffle passes) is too slow.
"""
from __future__ import annotations

import numpy as np


def _popcount(a):
    a = a.copy()
    c = np.zeros_like(a)
    while a.any():
        c += a & 1
        a >>= 1
    return c


class Lattice:
    """Lattice over K bits; bit a corresponds to event ev[a]; rate[a][s] = rate of adding
    bit a in state s (defined where bit a of s is 0); diag[s] = full diagonal of (D - Q)."""

    def __init__(self, K, rate, diag):
        self.K, self.rate, self.diag = K, rate, diag
        self.N = 1 << K
        s = np.arange(self.N)
        pc = _popcount(s)
        self.levels = [np.nonzero(pc == l)[0] for l in range(K + 1)]

    def solve(self, b, transpose=False):
        y = np.array(b, dtype=float)
        if not transpose:
            for lvl in self.levels:
                for a in range(self.K):
                    bit = 1 << a
                    sel = lvl[(lvl & bit) != 0]
                    y[sel] += self.rate[a][sel ^ bit] * y[sel ^ bit]
                y[lvl] /= self.diag[lvl]
        else:
            for lvl in reversed(self.levels):
                for a in range(self.K):
                    bit = 1 << a
                    sel = lvl[(lvl & bit) == 0]
                    y[sel] += self.rate[a][sel] * y[sel | bit]
                y[lvl] /= self.diag[lvl]
        return y


def _prod_table(w, bits_w):
    """table[u] = prod_{b in u} bits_w[b] for u in 0..2^len-1."""
    t = np.ones(1)
    for wb in bits_w:
        t = np.concatenate([t, t * wb])
    return t * w


def _event_rates(base, W, ev):
    """R[i][u] = base[i] * prod_{b in u, ev[b] != i} W[i][ev[b]] for all events i."""
    n_tot = len(base)
    return np.stack([_prod_table(base[i], [1.0 if e == i else W[i, e] for e in ev])
                     for i in range(n_tot)])


def _bit_sums(vals, K):
    """[sum over u with bit b set of vals[u]] for b in range(K), and the total."""
    u = np.arange(1 << K)
    return np.array([vals[(u >> b) & 1 == 1].sum() for b in range(K)]), vals.sum()


def _accumulate(G, i, ev, R_i, E_i, always=()):
    """G[i][ev[b]] += sum_{u has b, u lacks i} R_i E_i ; G[i][i] += total ; G[i][c] += total for c in always."""
    K = len(ev)
    u = np.arange(1 << K)
    w = R_i * E_i
    if i in ev:
        w = np.where((u >> ev.index(i)) & 1 == 1, 0.0, w)
    per_bit, tot = _bit_sums(w, K)
    for b in range(K):
        if ev[b] != i:
            G[i, ev[b]] += per_bit[b]
    G[i, i] += tot
    for c in always:
        G[i, c] += tot


def _single_space(base, W, ev, Dg, b, q_fn, extra_always=()):
    """Generic single-group space: solve (Dg + out - Q) y = b, adjoint with q = q_fn(y),
    return (y, x, G) where G[i][j] = d(q^T y)/d log W-ish parameters (rows = actor, cols = j)."""
    K = len(ev)
    n_tot = len(base)
    R = _event_rates(base, W, ev)                    # (n_tot, 2^K)
    u = np.arange(1 << K)
    member = np.zeros((n_tot, 1 << K), dtype=bool)
    for a, e in enumerate(ev):
        member[e] = (u >> a) & 1 == 1
    out = np.where(member, 0.0, R).sum(axis=0)
    lat = Lattice(K, [R[e] for e in ev], Dg + out)
    y = lat.solve(b)
    x = lat.solve(q_fn(y), transpose=True)
    G = np.zeros((n_tot, n_tot))
    g = x * y
    for i in range(n_tot):
        if R[i].max() == 0.0 and base[i] == 0.0:
            continue
        if i in ev:
            bit = 1 << ev.index(i)
            E = np.where(u & bit, 0.0, y * (x[u | bit] - x))
        else:
            E = -g
        _accumulate(G, i, list(ev), R[i], E, always=extra_always)
    return y, x, G, g


def prim_obs(log_theta, log_d_p, state_pt):
    """Types 0/1 (likelihood.py:387-405, 441-478).  Returns (logp, d_th, d_dp)."""
    n_tot = log_theta.shape[0]
    n = n_tot - 1
    ev = [j for j in range(n_tot) if state_pt[j]]
    if not ev:
        th = np.exp(np.diag(log_theta))
        return -np.log1p(th.sum()), np.diag(-th / (1.0 + th.sum())), np.zeros(n_tot)
    lt = log_theta.copy()
    lt[:-1, -1] = 0.0
    W = np.exp(lt - log_d_p[None, :])
    base = np.exp(np.diag(log_theta))
    K = len(ev)
    b = np.zeros(1 << K)
    b[0] = 1.0

    def q_fn(y):
        q = np.zeros_like(y)
        q[-1] = 1.0 / y[-1]
        return q

    y, x, G, _ = _single_space(base, W, ev, np.ones(1 << K), b, q_fn)
    off = G - np.diag(np.diag(G))
    d_dp = -off.sum(axis=0)
    d_th = G.copy()
    d_th[:-1, -1] = 0.0
    return np.log(y[-1]), d_th, d_dp


def met_obs(log_theta, log_d_p, log_d_m, state_mt):
    """Type 2 (likelihood.py:419-438, 481-512).  state_mt has n+1 entries, last = 1."""
    n_tot = log_theta.shape[0]
    n = n_tot - 1
    ev = [j for j in range(n_tot) if state_mt[j]]
    K = len(ev)
    u = np.arange(1 << K)
    seeded = (u >> (K - 1)) & 1 == 1
    dp, dm = np.exp(log_d_p), np.exp(log_d_m)
    Dp = _prod_table(1.0, [dp[e] for e in ev[:-1]] + [1.0])
    Dm = _prod_table(1.0, [dm[e] for e in ev])
    D = np.where(seeded, Dm, Dp)
    W = np.exp(log_theta)
    base = np.exp(np.diag(log_theta))
    b = np.zeros(1 << K)
    b[0] = 1.0

    def q_fn(y):
        q = np.zeros_like(y)
        q[-1] = 1.0 / y[-1]
        return q

    y, x, G, g = _single_space(base, W, ev, D, b, q_fn)
    d_dp, d_dm = np.zeros(n_tot), np.zeros(n_tot)
    for a, e in enumerate(ev):
        has = (u >> a) & 1 == 1
        d_dp[e] -= (g * D)[has & ~seeded].sum()
        d_dm[e] -= (g * D)[has & seeded].sum()
        d_dm[e] += 1.0
    return np.log(y[-1] * D[-1]), G, d_dp, d_dm


def coupled(order, log_theta, log_d_p, log_d_m, state_joint):
    """Type 3 (likelihood.py:286-384, 516-730; one_event.py).  order 0 / 1 / other.
    Returns (logp, d_th, d_dp, d_dm)."""
    n_tot = log_theta.shape[0]
    n = n_tot - 1
    pt = [int(state_joint[2 * j]) for j in range(n)]
    mt = [int(state_joint[2 * j + 1]) for j in range(n)]
    both = [j for j in range(n) if pt[j] and mt[j]]
    evA = both + [j for j in range(n) if pt[j] and not mt[j]]     # PT bits, shared events first
    evB = both + [j for j in range(n) if mt[j] and not pt[j]]     # MT bits, shared events first
    nb, KA, KB = len(both), len(evA), len(evB)
    NA, NB = 1 << KA, 1 << KB
    Th = np.exp(log_theta)
    base = np.exp(np.diag(log_theta))
    dp, dm = np.exp(log_d_p), np.exp(log_d_m)
    ev_idx = np.arange(n)

    # ---- pre-seeding diagonal sub-lattice over the shared events ---------------------------
    R0 = _event_rates(base, Th, both)                    # all n_tot events (row n = seeding)
    u0 = np.arange(1 << nb)
    mem0 = np.zeros((n_tot, 1 << nb), dtype=bool)
    for a, e in enumerate(both):
        mem0[e] = (u0 >> a) & 1 == 1
    DP0 = _prod_table(1.0, [dp[e] for e in both])
    diag0 = DP0 + np.where(mem0, 0.0, R0).sum(axis=0)
    lat0 = Lattice(nb, [R0[e] for e in both], diag0)
    e0 = np.zeros(1 << nb)
    e0[0] = 1.0
    y0 = lat0.solve(e0)

    # ---- post-seeding Kronecker-sum lattice, index s = uB << KA | uA ------------------------
    RP = _event_rates(base[:n], Th[:n, :n], evA)                              # (n, NA)
    RM = _event_rates(base[:n] * Th[:n, n], Th[:n, :n], evB)                  # (n, NB)
    uA, uB = np.arange(NA), np.arange(NB)
    memA = np.zeros((n, NA), dtype=bool)
    memB = np.zeros((n, NB), dtype=bool)
    for a, e in enumerate(evA):
        memA[e] = (uA >> a) & 1 == 1
    for a, e in enumerate(evB):
        memB[e] = (uB >> a) & 1 == 1
    DPA = _prod_table(dp[n], [dp[e] for e in evA])
    DMB = _prod_table(dm[n], [dm[e] for e in evB])
    dA = DPA + np.where(memA, 0.0, RP).sum(axis=0)
    dB = DMB + np.where(memB, 0.0, RM).sum(axis=0)
    diagJ = (dB[:, None] + dA[None, :]).ravel()
    rateJ = [np.broadcast_to(RP[e][None, :], (NB, NA)).ravel() for e in evA] + \
            [np.broadcast_to(RM[e][:, None], (NB, NA)).ravel() for e in evB]
    latJ = Lattice(KA + KB, rateJ, diagJ)
    emb = (u0 << KA) | u0                                 # shared events are the low bits of both groups
    bJ = np.zeros(NA * NB)
    bJ[emb] = R0[n] * y0
    y = latJ.solve(bJ)
    Y = y.reshape(NB, NA)

    # ---- second phases --------------------------------------------------------------------
    G = np.zeros((n_tot, n_tot))
    d_dp, d_dm = np.zeros(n_tot), np.zeros(n_tot)
    qJ = np.zeros((NB, NA))
    score = 0.0
    parts = []
    if order in (0, 1):                                   # PT observed first, MT evolves on
        cP = DPA[-1]
        W2 = np.exp(log_theta - log_d_m[None, :])
        R2 = _event_rates(base[:n] * W2[:n, n], W2[:n, :n], evB)
        lat2 = Lattice(KB, [R2[e] for e in evB], 1.0 + np.where(memB, 0.0, R2).sum(axis=0))
        v = cP * Y[:, NA - 1]
        p2 = lat2.solve(v)
        parts.append(("pf", p2[-1], lat2, R2, v, p2, cP))
        score += p2[-1]
    if order != 1:                                        # MT observed first, PT evolves on
        cM = DMB[-1]
        lt = log_theta.copy()
        lt[:-1, -1] = 0.0
        W3 = np.exp(lt - log_d_p[None, :])
        R3 = _event_rates(base[:n] * W3[:n, n], W3[:n, :n], evA)
        lat3 = Lattice(KA, [R3[e] for e in evA], 1.0 + np.where(memA, 0.0, R3).sum(axis=0))
        v = cM * Y[NB - 1, :]
        p3 = lat3.solve(v)
        parts.append(("mf", p3[-1], lat3, R3, v, p3, cM))
        score += p3[-1]
    for tag, _, lat2, R2, v, p2, c in parts:
        ev2 = evB if tag == "pf" else evA
        K2 = len(ev2)
        u2 = np.arange(1 << K2)
        q2 = np.zeros(1 << K2)
        q2[-1] = 1.0 / score
        x2 = lat2.solve(q2, transpose=True)
        G2 = np.zeros((n_tot, n_tot))
        for i in range(n):
            if i in ev2:
                bit = 1 << ev2.index(i)
                E = np.where(u2 & bit, 0.0, p2 * (x2[u2 | bit] - x2))
            else:
                E = -(x2 * p2)
            _accumulate(G2, i, list(ev2), R2[i], E, always=(n,))
        off = G2 - np.diag(np.diag(G2))
        t_direct = float(np.dot(x2, v))                   # d/d log c of x2 . (c * slice)
        if tag == "pf":
            G += G2
            d_dm -= off.sum(axis=0)
            qJ[:, NA - 1] += c * x2
            for e in evA + [n]:
                d_dp[e] += t_direct
        else:
            G2t = G2.copy()
            G2t[:-1, -1] = 0.0
            G += G2t
            d_dp -= off.sum(axis=0)
            qJ[NB - 1, :] += c * x2
            for e in evB + [n]:
                d_dm[e] += t_direct

    # ---- joint adjoint + marginal statistics ------------------------------------------------
    x = latJ.solve(qJ.ravel(), transpose=True)
    X = x.reshape(NB, NA)
    gJ = X * Y
    gA, gB = gJ.sum(axis=0), gJ.sum(axis=1)
    for i in range(n):
        if i in evA:
            bit = 1 << evA.index(i)
            Xs = X[:, uA | bit]
            E = np.where(uA & bit, 0.0, (Y * (Xs - X)).sum(axis=0))
        else:
            E = -gA
        _accumulate(G, i, list(evA), RP[i], E)
        if i in evB:
            bit = 1 << evB.index(i)
            Xs = X[uB | bit, :]
            E = np.where(uB & bit, 0.0, (Y * (Xs - X)).sum(axis=1))
        else:
            E = -gB
        _accumulate(G, i, list(evB), RM[i], E, always=(n,))
    for a, e in enumerate(evA):
        d_dp[e] -= (gA * DPA)[(uA >> a) & 1 == 1].sum()
    d_dp[n] -= (gA * DPA).sum()
    for a, e in enumerate(evB):
        d_dm[e] -= (gB * DMB)[(uB >> a) & 1 == 1].sum()
    d_dm[n] -= (gB * DMB).sum()

    # ---- pre-seeding adjoint ----------------------------------------------------------------
    xe = x[emb]
    x0 = lat0.solve(R0[n] * xe, transpose=True)
    _accumulate(G, n, list(both), R0[n], y0 * (xe - x0))
    for i in range(n):
        if i in both:
            bit = 1 << both.index(i)
            E = np.where(u0 & bit, 0.0, y0 * (x0[u0 | bit] - x0))
        else:
            E = -(x0 * y0)
        _accumulate(G, i, list(both), R0[i], E)
    g0 = x0 * y0 * DP0
    for a, e in enumerate(both):
        d_dp[e] -= g0[(u0 >> a) & 1 == 1].sum()
    return np.log(score), G, d_dp, d_dm


def patient_value_grad(log_theta, log_d_p, log_d_m, row):
    """Row dispatch as regularized_optimization.py:187-254.  Returns (is_type0, logp, g, dp, dm) or None."""
    n_tot = log_theta.shape[0]
    n = n_tot - 1
    typ = int(row[-1])
    z = np.zeros(n_tot)
    if typ in (0, 1):
        lp, g, d = prim_obs(log_theta, log_d_p, np.asarray(row[0:2 * n + 1:2]))
        return typ == 0, lp, g, d, z
    if typ == 2:
        st = np.append(np.asarray(row[1:2 * n:2]), 1)
        return (False,) + met_obs(log_theta, log_d_p, log_d_m, st)
    if typ == 3:
        order = int(row[-2])
        return (False,) + coupled(order if order in (0, 1) else 2, log_theta, log_d_p, log_d_m,
                                  np.asarray(row[0:2 * n + 1]))
    return None


def score_and_grad(log_theta, log_d_p, log_d_m, dat, perc_met):
    """Dataset level (regularized_optimization.py:163-267)."""
    n_tot = log_theta.shape[0]
    s = np.zeros(2)
    g = np.zeros((2, n_tot, n_tot))
    dpv, dmv = np.zeros((2, n_tot)), np.zeros((2, n_tot))
    for r in dat:
        out = patient_value_grad(log_theta, log_d_p, log_d_m, r)
        if out is None:
            continue
        k = 0 if out[0] else 1
        s[k] += out[1]
        g[k] += out[2]
        dpv[k] += out[3]
        dmv[k] += out[4]
    n_em = float(np.sum(dat[:, -3].astype(np.int64)))
    n_nm = dat.shape[0] - n_em
    w = perc_met * n_nm / ((1 - perc_met) * n_em) if n_em * n_nm != 0 else 1.0
    nf = w * n_em + n_nm
    return (w * s[1] + s[0]) / nf, (w * g[1] + g[0]) / nf, (w * dpv[1] + dpv[0]) / nf, w * dmv[1] / nf
