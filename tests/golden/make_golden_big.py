"""Golden vectors for the BIG-tier kernels: rows of the bench datasets themselves (SYN-v1, n = 25 / 100 000 patients and
n = 20 / 10 000 patients, BASELINE.json configs[3] / configs[2]) whose restricted lattices have 2^13 ... 2^18 states,
evaluated row by row with the UNMODIFIED reference (/root/reference) on the NumPy JAX shim.  Run in the build container:

    python tests/golden/make_golden_big.py         (about ten minutes)

Writes tests/golden/golden_big.npz: for every case the rows, the evaluation point and the reference's own
`score_and_grad` of each one-row dataset (`regularized_optimization.py:163`; with one row the weights are 1).
These are the sizes the tile kernels (`k_solve_tile`, `k_solve_tile_adjb`, `k_pf_*`) start at, so the GPU parity test on
this file checks them directly against reference-derived numbers, not through the second oracle.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_under_shim  # noqa: E402


def pick_rows(dat, n):
    typ = dat[:, -1]
    pt = dat[:, 0:2 * n:2].astype(int).sum(axis=1)
    mt = dat[:, 1:2 * n:2].astype(int).sum(axis=1)
    seed = dat[:, 2 * n].astype(int)
    order = dat[:, -2]
    rows = []
    # paired rows: lattice bits KA + KB (= k_J - 1) of 13, 15, 17; one row per order value where available
    for bits, orders in ((13, (0, 1, 2)), (15, (1, 2)), (17, (0,))):
        for o in orders:
            idx = np.nonzero((typ == 3) & (pt + mt == bits) & (order == o) & (pt >= 4) & (mt >= 3))[0]
            if idx.size:
                rows.append(int(idx[0]))
    # unpaired rows: PT-only (type 1) and MT-only (type 2) with 14, 16 and 18 lattice bits
    for k in (14, 16, 18):
        idx = np.nonzero((typ == 1) & (pt + seed == k))[0]
        if idx.size:
            rows.append(int(idx[0]))
        idx = np.nonzero((typ == 2) & (mt + 1 == k))[0]
        if idx.size:
            rows.append(int(idx[0]))
    return rows


def _one(args):
    n, ep, row = args
    regopt, lik, one, van, kv = ref_under_shim.load()
    import jax.numpy as jnp
    n_tot = n + 1
    th, dp, dm = ep[:n_tot * n_tot].reshape(n_tot, n_tot), ep[n_tot * n_tot:n_tot * (n_tot + 1)], ep[n_tot * (n_tot + 1):]
    t0 = time.time()
    s, a, b, c = regopt.score_and_grad(jnp.array(th), jnp.array(dp), jnp.array(dm), jnp.array(row[None, :]), 0.65)
    lp = float(np.asarray(s).reshape(-1)[0])
    print(f"n={n} type {row[-1]} order {row[-2]} bits {int(row[:2 * n + 1].sum())}: logp {lp:.12f} ({time.time() - t0:.1f} s)", flush=True)
    return lp, np.asarray(a), np.asarray(b), np.asarray(c)


def main():
    import multiprocessing as mp
    from metmhn_b200.simulate import syn_v1

    out, cases, jobs, meta = {}, [], [], []
    for n, n_dat in ((25, 100000), (20, 10000)):
        d = syn_v1(n, n_dat, 1000 * n + 3)                      # the dataset bench.py times
        dat, ep = d["dat"], d["eval_point"]
        rows = pick_rows(dat, n)
        if n == 20:
            rows = [r for r in rows if int(dat[r, :2 * n + 1].sum()) <= 16][:5]
        meta.append((n, n_dat, dat, ep, rows))
        jobs += [(n, ep, dat[r]) for r in rows]
    # the big rows first: the pool finishes when the longest one does
    order = sorted(range(len(jobs)), key=lambda i: -int(jobs[i][2][:-2].sum()) - (5 if jobs[i][2][-1] == 3 else 0))
    with mp.get_context("spawn").Pool(min(len(jobs), max(1, (os.cpu_count() or 2) - 2))) as pool:
        res_sorted = pool.map(_one, [jobs[i] for i in order], chunksize=1)
    res = [None] * len(jobs)
    for i, r in zip(order, res_sorted):
        res[i] = r
    pos = 0
    for n, n_dat, dat, ep, rows in meta:
        mine = res[pos:pos + len(rows)]
        pos += len(rows)
        name = f"bench_n{n}"
        for k, v in {"rows_index": np.asarray(rows), "rows": dat[rows], "eval_point": ep,
                     "row_logp": np.array([m[0] for m in mine]), "row_g": np.stack([m[1] for m in mine]),
                     "row_gdp": np.stack([m[2] for m in mine]), "row_gdm": np.stack([m[3] for m in mine]),
                     "seed": np.int64(1000 * n + 3), "n_dat": np.int64(n_dat)}.items():
            out[f"{name}/{k}"] = np.asarray(v)
        cases.append(name)
    out["cases"] = np.array(cases)
    path = os.path.join(HERE, "golden_big.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
