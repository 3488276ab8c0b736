"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on the
NumPy JAX shim (oracle/jax_shim).  Run in the build container only:

    python tests/golden/make_golden.py

Writes tests/golden/golden_v1.npz.  Every case stores its inputs and the outputs of the
reference's own `regularized_optimization.score`, `score_and_grad` and `score_and_grad_reg`
(`/root/reference/metmhn/regularized_optimization.py:55,163,270`), so the fixtures can be
checked on machines where the reference is absent (the GPU box).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_under_shim  # noqa: E402


def _params(n, rng, sd=0.6):
    n_tot = n + 1
    th = rng.normal(0.0, sd, (n_tot, n_tot))
    th[np.arange(n_tot), np.arange(n_tot)] = rng.normal(-1.0, 0.7, n_tot)
    return th, rng.normal(0.0, 0.4, n_tot), rng.normal(0.0, 0.4, n_tot)


def _row(n, rng, typ, order, p=0.5, empty=False):
    g = (rng.random(2 * n) < p).astype(np.int8)
    if empty:
        g[:] = 0
    seeding = 0 if typ == 0 else 1
    return np.concatenate([g, [seeding, order, typ]]).astype(np.int8)


def _kind_rows(n, rng, p):
    rows = [
        _row(n, rng, 0, -99, p), _row(n, rng, 0, -99, p, empty=True),
        _row(n, rng, 1, -99, p), _row(n, rng, 1, -99, p, empty=True),
        _row(n, rng, 2, -99, p), _row(n, rng, 2, -99, p, empty=True),
        _row(n, rng, 3, 0, p), _row(n, rng, 3, 1, p), _row(n, rng, 3, 2, p), _row(n, rng, 3, -99, p),
        _row(n, rng, 3, 0, p, empty=True), _row(n, rng, 3, 1, p, empty=True), _row(n, rng, 3, 2, p, empty=True),
    ]
    # a paired row with only shared events, one with only PT-private and one with only MT-private events
    r = _row(n, rng, 3, 0, p, empty=True); r[0:2] = 1; r[2:4] = 1; rows.append(r)
    r = _row(n, rng, 3, 1, p, empty=True); r[0] = 1; r[4] = 1; rows.append(r)
    r = _row(n, rng, 3, 2, p, empty=True); r[1] = 1; r[3] = 1; rows.append(r)
    return np.stack(rows)


def main():
    regopt, lik, one, van, kv = ref_under_shim.load()
    import jax.numpy as jnp

    out = {}
    cases = []

    def run_case(name, th, dp, dm, dat, perc_met, per_row, w_penal=None):
        jth, jdp, jdm, jdat = jnp.array(th), jnp.array(dp), jnp.array(dm), jnp.array(dat)
        rec = {"theta": th, "d_p": dp, "d_m": dm, "dat": dat, "perc_met": np.float64(perc_met)}
        if per_row:
            n_tot = th.shape[0]
            lp = np.zeros(dat.shape[0])
            lp_score = np.zeros(dat.shape[0])
            g = np.zeros((dat.shape[0], n_tot, n_tot))
            gdp = np.zeros((dat.shape[0], n_tot))
            gdm = np.zeros((dat.shape[0], n_tot))
            for r in range(dat.shape[0]):
                one_row = jdat[r:r + 1]
                s, a, b, c = regopt.score_and_grad(jth, jdp, jdm, one_row, perc_met)
                lp[r] = float(np.asarray(s).reshape(-1)[0])
                g[r], gdp[r], gdm[r] = np.asarray(a), np.asarray(b), np.asarray(c)
                lp_score[r] = float(np.asarray(regopt.score(jth, jdp, jdm, one_row, perc_met)).reshape(-1)[0])
            rec.update(row_logp=lp, row_logp_score=lp_score, row_g=g, row_gdp=gdp, row_gdm=gdm)
        s, a, b, c = regopt.score_and_grad(jth, jdp, jdm, jdat, perc_met)
        rec.update(score=np.float64(np.asarray(s).reshape(-1)[0]), g=np.asarray(a), gdp=np.asarray(b),
                   gdm=np.asarray(c),
                   score_only=np.float64(np.asarray(regopt.score(jth, jdp, jdm, jdat, perc_met)).reshape(-1)[0]))
        if w_penal is not None:
            params = np.concatenate([th.ravel(), dp, dm])
            f, gr = regopt.score_and_grad_reg(params, jdat, perc_met, regopt.symmetric_penal, w_penal)
            fr = regopt.score_reg(params, jdat, perc_met, regopt.symmetric_penal, w_penal)
            rec.update(w_penal=np.float64(w_penal), f_reg=np.float64(np.asarray(f).reshape(-1)[0]),
                       g_reg=np.asarray(gr).reshape(-1), f_reg_only=np.float64(np.asarray(fr).reshape(-1)[0]))
        for k, v in rec.items():
            out[f"{name}/{k}"] = np.asarray(v)
        cases.append(name)
        print(name, "rows", dat.shape[0], "score", rec["score"], flush=True)

    # per-kind rows at several n (every patient kind, empty / ragged genotypes, unknown order marker)
    for n, seed, p in [(1, 101, 0.6), (2, 102, 0.6), (3, 103, 0.6), (4, 104, 0.55), (5, 105, 0.5), (6, 106, 0.45)]:
        rng = np.random.default_rng(seed)
        th, dp, dm = _params(n, rng)
        dat = _kind_rows(n, rng, p) if n >= 2 else np.stack(
            [_row(n, rng, t, o, p) for t, o in [(0, -99), (1, -99), (2, -99), (3, 0), (3, 1), (3, 2)]]
            + [_row(n, rng, 3, o, p, empty=True) for o in (0, 1, 2)] + [_row(n, rng, 0, -99, p, empty=True)])
        run_case(f"kinds_n{n}", th, dp, dm, dat, 0.65, per_row=True, w_penal=0.05)

    # mixed dataset with weighting and penalty, two perc_met values
    from metmhn_b200.simulate import syn_v1
    d = syn_v1(5, 60, 5005)
    th, dp, dm = d["theta"] + 0.0, d["d_p"], d["d_m"]
    run_case("mixed_n5_pm65", th, dp, dm, d["dat"], 0.65, per_row=False, w_penal=0.01)
    rng = np.random.default_rng(7)
    th2, dp2, dm2 = _params(5, rng, sd=0.4)
    run_case("mixed_n5_pm20", th2, dp2, dm2, d["dat"], 0.2, per_row=False, w_penal=0.4)
    # rows of an unknown type contribute nothing but are counted (Appendix B.9)
    dat_u = d["dat"][:20].copy()
    dat_u[3, -1] = 7
    run_case("unknown_type_n5", th2, dp2, dm2, dat_u, 0.65, per_row=False)
    # only type-0 rows -> weight w = 1 branch
    dat0 = d["dat"][d["dat"][:, -1] == 0][:6]
    if dat0.shape[0]:
        run_case("only_type0_n5", th2, dp2, dm2, dat0, 0.65, per_row=False)

    # larger restricted spaces: n = 8 (joint up to 2^11) and a few rows at n = 10
    d8 = syn_v1(8, 40, 8008)
    run_case("syn_n8", d8["theta"], d8["d_p"], d8["d_m"], d8["dat"], 0.65, per_row=True, w_penal=0.001)
    d10 = syn_v1(10, 400, 10010)
    dat = d10["dat"]
    kj = dat[:, :21].sum(axis=1)
    pick = np.concatenate([np.nonzero((dat[:, -1] == 3) & (kj >= 9) & (kj <= 12))[0][:4],
                           np.nonzero((dat[:, -1] == 2) & (kj >= 7))[0][:2],
                           np.nonzero((dat[:, -1] == 1) & (kj >= 7))[0][:2]])
    ep = d10["eval_point"]
    run_case("syn_n10", ep[:121].reshape(11, 11), ep[121:132], ep[132:], dat[pick], 0.65, per_row=True)

    out["cases"] = np.array(cases)
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
