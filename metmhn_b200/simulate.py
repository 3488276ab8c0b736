"""Synthetic metMHN data: a vectorised NumPy Gillespie sampler and the SYN-v1 benchmark
datasets.

The sampler restates the process of the reference's `metmhn/simulations.py:8-77`
(`single_traject`): primary tumour (PT) and metastasis (MT) evolve in lock-step until
the seeding event, independently afterwards; a run stops when the PT is diagnosed before
seeding or when both tumours are diagnosed.  Row layout as the likelihood expects it
(`regularized_optimization.py:63-66`): [PT_0, MT_0, ..., PT_{n-1}, MT_{n-1}, seeding,
order, type].
"""
from __future__ import annotations

import numpy as np


def random_params(n: int, rng: np.random.Generator, density: float = 0.2,
                  off_sd: float = 0.75, d_sd: float = 0.3):
    """Ground-truth parameters of SYN-v1: decaying base rates, sparse normal interactions."""
    n_tot = n + 1
    th = np.zeros((n_tot, n_tot))
    mask = rng.random((n_tot, n_tot)) < density
    th[mask] = rng.normal(0.0, off_sd, size=int(mask.sum()))
    diag = np.zeros(n_tot)
    diag[:n] = -0.5 - 2.5 * np.arange(n) / max(n - 1, 1)
    th[np.arange(n_tot), np.arange(n_tot)] = diag
    return th, rng.normal(0.0, d_sd, n_tot), rng.normal(0.0, d_sd, n_tot)


def simulate(log_theta, log_d_p, log_d_m, n_sim: int, rng: np.random.Generator):
    """Sample n_sim trajectories.  Returns (geno int8 (n_sim, 2n+1), order int8 (n_sim,)):
    order is 1 (PT diagnosed first) / 2 (MT first) for runs that seeded, 0 otherwise."""
    n_tot = log_theta.shape[0]
    base = np.diagonal(log_theta)
    th_pt = log_theta.copy()
    th_pt[:-1, -1] = 0.0                      # seeding does not act on the PT (simulations.py:62-64)
    pt = np.zeros((n_sim, n_tot), dtype=np.float64)
    mt = np.zeros((n_sim, n_tot), dtype=np.float64)
    obs_pt = np.zeros(n_sim, dtype=bool)
    obs_mt = np.zeros(n_sim, dtype=bool)
    order = np.zeros(n_sim, dtype=np.int8)
    alive = np.ones(n_sim, dtype=bool)
    while alive.any():
        idx = np.nonzero(alive)[0]
        p, m = pt[idx], mt[idx]
        seeded = p[:, -1] > 0
        r_pt = np.exp(p @ th_pt.T + base) * (1.0 - p)
        r_pt_obs = np.exp(p @ log_d_p)
        pt_live = ~obs_pt[idx]
        r_pt *= pt_live[:, None]
        r_pt_obs = r_pt_obs * pt_live
        mt_live = seeded & ~obs_mt[idx]
        r_mt = np.exp(m @ log_theta.T + base) * (1.0 - m) * mt_live[:, None]
        r_mt_obs = np.exp(m @ log_d_m) * mt_live
        rates = np.concatenate([r_pt, r_pt_obs[:, None], r_mt, r_mt_obs[:, None]], axis=1)
        cum = np.cumsum(rates, axis=1)
        u = rng.random(idx.shape[0]) * cum[:, -1]
        ev = np.minimum((cum <= u[:, None]).sum(axis=1), rates.shape[1] - 1)
        is_pt_mut = ev < n_tot
        is_pt_obs = ev == n_tot
        is_mt_mut = (ev > n_tot) & (ev < 2 * n_tot + 1)
        is_mt_obs = ev == 2 * n_tot + 1
        # PT-side events before seeding hit both copies (simulations.py:52-55)
        rows = idx[is_pt_mut]
        cols = ev[is_pt_mut]
        pt[rows, cols] = 1.0
        both = ~seeded[is_pt_mut]
        mt[rows[both], cols[both]] = 1.0
        rows = idx[is_pt_obs]
        first = ~obs_mt[rows] & seeded[is_pt_obs]
        order[rows[first]] = 1
        obs_pt[rows] = True
        obs_mt[rows[~seeded[is_pt_obs]]] = True
        rows = idx[is_mt_mut]
        mt[rows, ev[is_mt_mut] - n_tot - 1] = 1.0
        rows = idx[is_mt_obs]
        order[rows[~obs_pt[rows]]] = 2
        obs_mt[rows] = True
        seeded_now = pt[idx, -1] > 0
        alive[idx] = ~((obs_pt[idx] & obs_mt[idx]) | (obs_pt[idx] & ~seeded_now))
    geno = np.empty((n_sim, 2 * n_tot - 1), dtype=np.int8)
    geno[:, 0:-1:2] = pt[:, :-1]
    geno[:, 1::2] = mt[:, :-1]
    geno[:, -1] = pt[:, -1]
    return geno, order


def simulate_gpu(log_theta, log_d_p, log_d_m, n_sim: int, seed: int = 0, device: int = 0):
    """The same process on the GPU (mmh_simulate, one thread per trajectory): (geno, order) in the layout of `simulate`.
    Different random stream than the NumPy sampler, same distribution (tests/test_gpu_workloads.py)."""
    from ._lib import simulate as _sim
    return _sim(log_theta, log_d_p, log_d_m, n_sim, seed, device)


def syn_v1(n: int, n_dat: int, seed: int, max_joint_bits: int = 24, heavy: bool = False):
    """SYN-v1 dataset.  Returns dict(dat, theta, d_p, d_m, eval_point, perc_met).

    Composition follows `examples/recall_study.py:120-125`: 11.5 % never-metastasised PTs
    (type 0, all-zero genotypes kept), of the rest 10.7 % paired (type 3, a quarter of them
    with unknown order 0), 38.6 % PT-only (type 1), remainder MT-only (type 2).
    Paired rows whose joint state has more than max_joint_bits set bits are redrawn."""
    rng = np.random.Generator(np.random.PCG64(seed))
    th, dp, dm = random_params(n, rng)
    if heavy:
        th[np.arange(n), np.arange(n)] += 1.0
    n0 = int(round(0.115 * n_dat))
    n_em = n_dat - n0
    n3 = int(round(0.107 * n_em))
    n1 = int(round(0.386 * n_em))
    n2 = n_em - n3 - n1
    rows0, rows_em = [], []
    have0 = have_em = 0
    while have0 < n0 or have_em < n_em:
        g, o = simulate(th, dp, dm, max(4096, n_dat), rng)
        sd = g[:, -1] == 1
        rows0.append(g[~sd])
        rows_em.append(np.concatenate([g[sd], o[sd, None]], axis=1))
        have0 += int((~sd).sum())
        have_em += int(sd.sum())
    g0 = np.concatenate(rows0)[:n0]
    gem = np.concatenate(rows_em)
    dat = np.zeros((n_dat, 2 * n + 3), dtype=np.int8)
    dat[:n0, :2 * n + 1] = g0
    dat[:n0, 1:2 * n:2] = 0
    dat[:n0, -2] = -99
    dat[:n0, -1] = 0
    # paired rows: take from the metastasised pool, skipping over-large joint states
    kj = gem[:, :2 * n + 1].sum(axis=1)
    ok = np.nonzero(kj <= max_joint_bits)[0]
    redrawn = int((kj[: n3] > max_joint_bits).sum())
    pick3 = ok[:n3]
    rest = np.setdiff1d(np.arange(gem.shape[0]), pick3, assume_unique=True)[: n1 + n2]
    r = n0
    dat[r:r + n3, :2 * n + 1] = gem[pick3, :2 * n + 1]
    dat[r:r + n3, -2] = gem[pick3, -1]
    unknown = rng.random(n3) < 0.25
    dat[r:r + n3, -2][unknown] = 0
    dat[r:r + n3, -1] = 3
    r += n3
    dat[r:r + n1, :2 * n + 1] = gem[rest[:n1], :2 * n + 1]
    dat[r:r + n1, 1:2 * n:2] = 0
    dat[r:r + n1, -2] = -99
    dat[r:r + n1, -1] = 1
    r += n1
    dat[r:, :2 * n + 1] = gem[rest[n1:n1 + n2], :2 * n + 1]
    dat[r:, 0:2 * n:2] = 0
    dat[r:, -2] = -99
    dat[r:, -1] = 2
    dat = dat[rng.permutation(n_dat)]
    npar = (n + 1) * (n + 3)
    eval_point = np.concatenate([th.ravel(), dp, dm]) + rng.normal(0.0, 0.1, npar)
    return dict(dat=np.ascontiguousarray(dat), theta=th, d_p=dp, d_m=dm,
                eval_point=eval_point, perc_met=0.65, redrawn=redrawn)


def k_histogram(dat: np.ndarray):
    """Histogram of restricted state-space sizes per patient type: {type: {k: count}}."""
    n = (dat.shape[1] - 3) // 2
    out = {}
    for t in (0, 1, 2, 3):
        sub = dat[dat[:, -1] == t]
        if t in (0, 1):
            k = sub[:, 0:2 * n + 1:2].sum(axis=1)
        elif t == 2:
            k = sub[:, 1:2 * n:2].sum(axis=1) + 1
        else:
            k = sub[:, :2 * n + 1].sum(axis=1)
        vals, cnt = np.unique(k, return_counts=True)
        out[t] = {int(v): int(c) for v, c in zip(vals, cnt)}
    return out
