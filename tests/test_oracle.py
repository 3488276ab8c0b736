"""CPU tests: the two oracle formulations against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), and against each other."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import lattice_direct as ld
from oracle import reference_restated as rr

TOL = 1e-10      # north_star tolerance (relative, FP64)


def _case(golden, c):
    return tuple(golden[f"{c}/{k}"] for k in ("theta", "d_p", "d_m", "dat")) + (float(golden[f"{c}/perc_met"]),)


def _cases(golden):
    return [str(c) for c in golden["cases"]]


@pytest.mark.parametrize("mod", [rr, ld], ids=["restated", "lattice"])
def test_dataset_level_matches_reference(golden, mod):
    for c in _cases(golden):
        th, dp, dm, dat, pm = _case(golden, c)
        s, g, a, b = mod.score_and_grad(th, dp, dm, dat, pm)
        assert abs(s - golden[f"{c}/score"]) <= TOL * abs(golden[f"{c}/score"]), c
        assert rel_err(g, golden[f"{c}/g"]) <= TOL, c
        assert rel_err(a, golden[f"{c}/gdp"]) <= TOL, c
        assert rel_err(b, golden[f"{c}/gdm"]) <= TOL, c


@pytest.mark.parametrize("mod", [rr, ld], ids=["restated", "lattice"])
def test_every_patient_kind_matches_reference(golden, mod):
    for c in _cases(golden):
        if f"{c}/row_logp" not in golden.files:
            continue
        th, dp, dm, dat, _ = _case(golden, c)
        for r in range(dat.shape[0]):
            _, lp, g, a, b = mod.patient_value_grad(th, dp, dm, dat[r])
            assert abs(lp - golden[f"{c}/row_logp"][r]) <= TOL * abs(golden[f"{c}/row_logp"][r]), (c, r)
            for got, key in ((g, "row_g"), (a, "row_gdp"), (b, "row_gdm")):
                ref = golden[f"{c}/{key}"][r]
                if np.abs(ref).max() == 0.0:
                    assert np.abs(got).max() == 0.0, (c, r, key)       # structural zeros stay exact
                else:
                    assert rel_err(got, ref) <= TOL, (c, r, key)


def test_value_only_and_penalised_objective(golden):
    for c in _cases(golden):
        th, dp, dm, dat, pm = _case(golden, c)
        assert abs(rr.score(th, dp, dm, dat, pm) - golden[f"{c}/score_only"]) <= TOL * abs(golden[f"{c}/score_only"])
        if f"{c}/f_reg" in golden.files:
            params = np.concatenate([th.ravel(), dp, dm])
            lam = float(golden[f"{c}/w_penal"])
            f, gr = rr.score_and_grad_reg(params, dat, pm, rr.symmetric_penal, lam)
            assert abs(f - golden[f"{c}/f_reg"]) <= TOL * abs(golden[f"{c}/f_reg"])
            assert rel_err(gr, golden[f"{c}/g_reg"]) <= TOL
            assert abs(rr.score_reg(params, dat, pm, rr.symmetric_penal, lam) - golden[f"{c}/f_reg_only"]) <= 1e-12


def test_score_equals_score_and_grad_value(golden):
    """Appendix B.8 of SURVEY.md: the reference's FD tests rely on this."""
    th, dp, dm, dat, pm = _case(golden, "mixed_n5_pm65")
    assert abs(rr.score(th, dp, dm, dat, pm) - rr.score_and_grad(th, dp, dm, dat, pm)[0]) < 1e-13


def test_gradient_is_the_derivative_by_finite_differences():
    """The reference's own test idea (tests/test_gradient.py:71-157): analytic vs forward FD."""
    rng = np.random.default_rng(3)
    n = 3
    th = rng.normal(0, 0.5, (n + 1, n + 1))
    dp, dm = np.log([1.0, 2.0, 3.0, 4.0]), np.log([0.5, 1.5, 2.5, 3.5])
    rows = [[1, 0, 1, 1, 0, 1, 1, 0, 3], [1, 1, 0, 1, 1, 0, 1, 1, 3], [0, 1, 1, 1, 0, 0, 1, 2, 3],
            [1, 0, 1, 0, 0, 0, 1, -99, 1], [0, 1, 0, 1, 0, 1, 1, -99, 2], [1, 0, 0, 0, 1, 0, 0, -99, 0]]
    dat = np.array(rows, dtype=np.int8)
    s, g, a, b = ld.score_and_grad(th, dp, dm, dat, 0.8)
    h = 1e-7
    for i in range(n + 1):
        for j in range(n + 1):
            t2 = th.copy(); t2[i, j] += h
            fd = (ld.score_and_grad(t2, dp, dm, dat, 0.8)[0] - s) / h
            assert abs(fd - g[i, j]) < 1e-5
        d2 = dp.copy(); d2[i] += h
        assert abs((ld.score_and_grad(th, d2, dm, dat, 0.8)[0] - s) / h - a[i]) < 1e-5
        d2 = dm.copy(); d2[i] += h
        assert abs((ld.score_and_grad(th, dp, d2, dat, 0.8)[0] - s) / h - b[i]) < 1e-5


def test_two_formulations_agree_on_mid_size_spaces():
    """Sizes beyond the golden fixtures: n = 12, joint spaces up to 2^12."""
    from metmhn_b200.simulate import syn_v1
    d = syn_v1(12, 300, 12012)
    dat = d["dat"]
    kj = dat[:, :25].sum(axis=1)
    pick = np.concatenate([np.nonzero((dat[:, -1] == 3) & (kj >= 10) & (kj <= 13))[0][:3],
                           np.nonzero((dat[:, -1] == 2) & (kj >= 8))[0][:2],
                           np.nonzero((dat[:, -1] == 1) & (kj >= 8))[0][:2]])
    ep = d["eval_point"]
    th, dp, dm = ep[:169].reshape(13, 13), ep[169:182], ep[182:]
    for r in pick:
        o1 = rr.patient_value_grad(th, dp, dm, dat[r])
        o2 = ld.patient_value_grad(th, dp, dm, dat[r])
        assert abs(o1[1] - o2[1]) <= TOL * abs(o1[1])
        for x, y in zip(o1[2:], o2[2:]):
            assert rel_err(y, x) <= TOL


def test_lattice_oracle_matches_reference_on_big_lattices():
    """tests/golden/golden_big.npz: rows of the bench datasets with 2^13 ... 2^18-state lattices evaluated by the UNMODIFIED
    reference under the shim (tests/golden/make_golden_big.py, hours of CPU).  The lattice formulation -- the checker of
    the big-tier GPU tests -- reproduces them, so the oracle chain is pinned to the reference where the time goes."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_big.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/golden_big.npz has not been generated")
    gb = np.load(path)
    checked = 0
    for c in [str(x) for x in gb["cases"]]:
        rows, ep = gb[f"{c}/rows"], gb[f"{c}/eval_point"]
        n_tot = (rows.shape[1] - 3) // 2 + 1
        th, dp, dm = ep[:n_tot * n_tot].reshape(n_tot, n_tot), ep[n_tot * n_tot:n_tot * n_tot + n_tot], ep[n_tot * n_tot + n_tot:]
        bits = rows[:, :2 * (n_tot - 1) + 1].astype(int).sum(axis=1)
        for r in np.nonzero(bits <= 16)[0]:                      # seconds in total; the 2^18-state rows are left to the GPU test
            out = ld.patient_value_grad(th, dp, dm, rows[r])
            ref = gb[f"{c}/row_logp"][r]
            assert abs(out[1] - ref) <= TOL * abs(ref), (c, r)
            assert rel_err(out[2], gb[f"{c}/row_g"][r]) <= TOL, (c, r)
            assert rel_err(out[3], gb[f"{c}/row_gdp"][r]) <= TOL, (c, r)
            assert rel_err(out[4], gb[f"{c}/row_gdm"][r]) <= TOL, (c, r)
            checked += 1
    assert checked >= 10
