"""GPU parity tests: the CUDA path (through the C-ABI) against the golden vectors of the
unmodified reference and against the CPU oracle on seeded synthetic data."""
import numpy as np
import pytest

from conftest import rel_err

TOL = 1e-10          # relative, FP64 (BASELINE.json north_star)

pytestmark = pytest.mark.gpu


def _split(g, n_tot):
    sq = n_tot * n_tot
    return g[:sq].reshape(n_tot, n_tot), g[sq:sq + n_tot], g[sq + n_tot:]


def _case(golden, c):
    return tuple(golden[f"{c}/{k}"] for k in ("theta", "d_p", "d_m", "dat")) + (float(golden[f"{c}/perc_met"]),)


def test_dataset_level_against_reference_golden(golden):
    import metmhn_b200 as mm
    for c in [str(x) for x in golden["cases"]]:
        th, dp, dm, dat, pm = _case(golden, c)
        s, g, a, b = mm.score_and_grad(th, dp, dm, dat, pm)
        ref = golden[f"{c}/score"]
        assert abs(s - ref) <= TOL * abs(ref), (c, s, ref)
        assert rel_err(g, golden[f"{c}/g"]) <= TOL, c
        assert rel_err(a, golden[f"{c}/gdp"]) <= TOL, c
        assert rel_err(b, golden[f"{c}/gdm"]) <= TOL, c
        assert abs(mm.score(th, dp, dm, dat, pm) - golden[f"{c}/score_only"]) <= TOL * abs(ref), c
        if f"{c}/f_reg" in golden.files:
            params = np.concatenate([th.ravel(), dp, dm])
            lam = float(golden[f"{c}/w_penal"])
            f, gr = mm.score_and_grad_reg(params, dat, pm, mm.symmetric_penal, lam)
            assert abs(float(f) - golden[f"{c}/f_reg"]) <= TOL * abs(golden[f"{c}/f_reg"]), c
            assert rel_err(gr, golden[f"{c}/g_reg"]) <= TOL, c
            assert abs(float(mm.score_reg(params, dat, pm, mm.symmetric_penal, lam)) - golden[f"{c}/f_reg_only"]) <= 1e-10


def test_every_patient_kind_against_reference_golden(golden):
    """One handle per row: every kind, empty / ragged genotypes, the seeding-only paired rows that the
    reference routes through one_event.py, the -99 order marker."""
    from metmhn_b200 import Handle
    for c in [str(x) for x in golden["cases"]]:
        if f"{c}/row_logp" not in golden.files:
            continue
        th, dp, dm, dat, _ = _case(golden, c)
        params = np.concatenate([th.ravel(), dp, dm])
        n_tot = th.shape[0]
        hall = Handle(dat)
        lp_all = hall.per_patient(params)
        lp_g, g_all = hall.per_patient_grads(params)                 # mmh_per_patient_grads: same numbers as one handle per row
        for r in range(dat.shape[0]):
            h = Handle(dat[r:r + 1])
            s, g = h.eval_weighted(params, 1.0, 1.0)
            h.close()
            ref = golden[f"{c}/row_logp"][r]
            assert abs(s - ref) <= TOL * abs(ref), (c, r)
            assert abs(lp_all[r] - ref) <= TOL * abs(ref), (c, r)
            assert lp_g[r] == s and np.array_equal(g_all[r], g), (c, r)
            for got, key in zip(_split(g, n_tot), ("row_g", "row_gdp", "row_gdm")):
                want = golden[f"{c}/{key}"][r]
                if np.abs(want).max() == 0.0:
                    assert np.abs(got).max() == 0.0, (c, r, key)
                else:
                    assert rel_err(got, want) <= TOL, (c, r, key)


def test_synthetic_n10_against_oracle():
    """BASELINE config 2 shape (n = 10, mixed paired / unpaired) at a size the oracle finishes quickly."""
    import metmhn_b200 as mm
    from metmhn_b200.simulate import syn_v1
    from oracle import lattice_direct as ld
    d = syn_v1(10, 300, 10001, max_joint_bits=15)
    ep = d["eval_point"]
    th, dp, dm = ep[:121].reshape(11, 11), ep[121:132], ep[132:]
    s, g, a, b = mm.score_and_grad(th, dp, dm, d["dat"], 0.65)
    s0, g0, a0, b0 = ld.score_and_grad(th, dp, dm, d["dat"], 0.65)
    assert abs(s - s0) <= TOL * abs(s0)
    assert rel_err(g, g0) <= TOL and rel_err(a, a0) <= TOL and rel_err(b, b0) <= TOL


def test_big_tier_spaces_against_oracle():
    """Spaces above the small/big tier boundary (level launches): n = 14, paired rows with 13..17 bits."""
    from metmhn_b200 import Handle
    from metmhn_b200.simulate import syn_v1
    from oracle import lattice_direct as ld
    d = syn_v1(14, 1500, 14014, max_joint_bits=18)
    dat = d["dat"]
    kj = dat[:, :29].sum(axis=1)
    pick = np.concatenate([np.nonzero((dat[:, -1] == 3) & (kj >= 14) & (kj <= 18))[0][:5],
                           np.nonzero((dat[:, -1] == 2) & (kj >= 12))[0][:2],
                           np.nonzero((dat[:, -1] == 1) & (kj >= 13))[0][:2]])
    assert len(pick) >= 5
    ep = d["eval_point"]
    th, dp, dm = ep[:225].reshape(15, 15), ep[225:240], ep[240:]
    sub = np.ascontiguousarray(dat[pick])
    h = Handle(sub)
    lp = h.per_patient(ep)
    s, g = h.eval_weighted(ep, 1.0, 1.0)
    G = np.zeros(15 * 17)
    S = 0.0
    for r in range(sub.shape[0]):
        out = ld.patient_value_grad(th, dp, dm, sub[r])
        assert abs(lp[r] - out[1]) <= TOL * abs(out[1]), r
        S += out[1]
        G += np.concatenate([out[2].ravel(), out[3], out[4]])
    assert abs(s - S) <= TOL * abs(S)
    assert rel_err(g, G) <= TOL


def test_wide_groups_against_oracle():
    """Groups with more than 16 bits use product tables: unpaired patients with 17 events and a lopsided pair."""
    from metmhn_b200 import Handle
    from oracle import lattice_direct as ld
    n = 18
    rng = np.random.default_rng(18)
    th = rng.normal(0.0, 0.3, (n + 1, n + 1))
    th[np.arange(n + 1), np.arange(n + 1)] = rng.normal(-1.0, 0.5, n + 1)
    dp, dm = rng.normal(0, 0.3, n + 1), rng.normal(0, 0.3, n + 1)
    rows = np.zeros((4, 2 * n + 3), dtype=np.int8)
    rows[0, 0:34:2] = 1; rows[0, 2 * n] = 1; rows[0, -2:] = (-99, 1)            # type 1, 17 PT events + seeding
    rows[1, 1:34:2] = 1; rows[1, 2 * n] = 1; rows[1, -2:] = (-99, 2)            # type 2, 17 MT events
    rows[2, 0:34:2] = 1; rows[2, 1] = 1; rows[2, 35] = 1; rows[2, 2 * n] = 1; rows[2, -2:] = (0, 3)
    rows[3, 1:34:2] = 1; rows[3, 0] = 1; rows[3, 34] = 1; rows[3, 2 * n] = 1; rows[3, -2:] = (2, 3)
    params = np.concatenate([th.ravel(), dp, dm])
    for r in range(rows.shape[0]):
        h = Handle(rows[r:r + 1])
        s, g = h.eval_weighted(params, 1.0, 1.0)
        h.close()
        out = ld.patient_value_grad(th, dp, dm, rows[r])
        assert abs(s - out[1]) <= TOL * abs(out[1]), r
        assert rel_err(g, np.concatenate([out[2].ravel(), out[3], out[4]])) <= TOL, r


def test_chunking_and_repeatability():
    """Small scratch chunks give the same numbers; repeated evaluations are bit-identical."""
    from metmhn_b200 import Handle
    from metmhn_b200.simulate import syn_v1
    d = syn_v1(8, 400, 8123)
    ep = d["eval_point"]
    h1 = Handle(d["dat"])
    h2 = Handle(d["dat"], chunk_bytes=1 << 16)
    s1, g1 = h1.value_grad(ep, 0.65)
    s1b, g1b = h1.value_grad(ep, 0.65)
    s2, g2 = h2.value_grad(ep, 0.65)
    assert s1 == s1b and np.array_equal(g1, g1b)
    assert h2.stats()["n_chunks"] > h1.stats()["n_chunks"]
    assert abs(s1 - s2) <= 1e-13 * abs(s1) and rel_err(g2, g1) <= 1e-12


def test_error_paths():
    from metmhn_b200 import Handle, MetMHNError
    bad = np.zeros((1, 9), dtype=np.int8)
    bad[0, -1] = 3                      # paired row without seeding
    with pytest.raises(MetMHNError):
        Handle(bad)
    bad2 = np.zeros((1, 9), dtype=np.int8)
    bad2[0, 0] = 2
    with pytest.raises(MetMHNError):
        Handle(bad2)
    h = Handle(np.zeros((0, 9), dtype=np.int8))       # empty dataset is allowed
    assert h.stats()["n_spaces"] == 0


def _pair_row(n, pt, mt, order):
    row = np.zeros(2 * n + 3, dtype=np.int8)
    for e in pt:
        row[2 * e] = 1
    for e in mt:
        row[2 * e + 1] = 1
    row[2 * n] = 1
    row[-2:] = (order, 3)
    return row


def test_tile_kernel_shapes_against_oracle():
    """Every shape class of the big-tier solve: column blocks below / above 32 per level (fused group-B
    statistics with several row groups or several column chunks per CTA), few rows, the generic kernel for pairs
    with fewer than 4 PT bits, shared and private events on both sides."""
    from metmhn_b200 import Handle
    from oracle import lattice_direct as ld
    n = 16
    rng = np.random.default_rng(1616)
    th = rng.normal(0.0, 0.4, (n + 1, n + 1))
    th[np.arange(n + 1), np.arange(n + 1)] = rng.normal(-1.0, 0.5, n + 1)
    dp, dm = rng.normal(0, 0.3, n + 1), rng.normal(0, 0.3, n + 1)
    ev = list(range(n))
    rows = np.stack([
        _pair_row(n, ev[:12], ev[10:14], 1),          # KA = 12, KB = 4: 70 column blocks at the middle level
        _pair_row(n, ev[:11], ev[8:13], 0),           # KA = 11, KB = 5, both second phases
        _pair_row(n, ev[:5], ev[3:13], 2),            # KA = 5,  KB = 10: two column blocks, 16 row groups per CTA
        _pair_row(n, ev[:3], ev[2:14], 1),            # KA = 3: generic kernel
        _pair_row(n, ev[:13], ev[13:15], 0),          # KA = 13, KB = 2: few rows, no shared event
        _pair_row(n, ev[:8], ev[:8], 1),              # identical tumours
    ])
    params = np.concatenate([th.ravel(), dp, dm])
    for r in range(rows.shape[0]):
        h = Handle(rows[r:r + 1])
        s, g = h.eval_weighted(params, 1.0, 1.0)
        s_again, g_again = h.eval_weighted(params, 1.0, 1.0)
        h.close()
        assert s == s_again and np.array_equal(g, g_again), r
        out = ld.patient_value_grad(th, dp, dm, rows[r])
        assert abs(s - out[1]) <= TOL * abs(out[1]), r
        assert rel_err(g, np.concatenate([out[2].ravel(), out[3], out[4]])) <= TOL, r


def test_alternative_kernel_paths_agree():
    """MMH_ROWBLK=1 sends the pairs to the shared-memory row-block kernel (k_solve_rb, mmh_rowblock.cuh) instead of the
    tile kernels (k_solve_tile / k_solve_tile_adjb).  Both paths give the same numbers (the switch is read once per
    process: run each in a fresh interpreter)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from metmhn_b200 import Handle\n"
            "from metmhn_b200.simulate import syn_v1\n"
            "d = syn_v1(14, 1500, 14014, max_joint_bits=18)\n"
            "s, g = Handle(d['dat']).value_grad(d['eval_point'], 0.65)\n"
            "print(repr(float(s))); print(' '.join(repr(float(v)) for v in g))\n") % root
    res = {}
    for name, env in (("default", {}), ("rowblock", {"MMH_ROWBLK": "1"})):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = out.stdout.strip().splitlines()
        res[name] = (float(lines[-2]), np.array([float(v) for v in lines[-1].split()]))
    s0, g0 = res["default"]
    s, g = res["rowblock"]
    assert abs(s - s0) <= 1e-12 * abs(s0)
    assert rel_err(g, g0) <= 1e-11


def test_wide_pairs_against_oracle():
    """Pairs with a group of more than 16 bits (product tables, generic solve kernel): wide PT side with shared events,
    wide MT side, against the lattice oracle."""
    from metmhn_b200 import Handle
    from oracle import lattice_direct as ld
    n = 21
    rng = np.random.default_rng(2121)
    th = rng.normal(0.0, 0.3, (n + 1, n + 1))
    th[np.arange(n + 1), np.arange(n + 1)] = rng.normal(-1.0, 0.5, n + 1)
    dp, dm = rng.normal(0, 0.3, n + 1), rng.normal(0, 0.3, n + 1)
    ev = list(range(n))
    rows = np.stack([
        _pair_row(n, ev[:17], ev[15:18], 1),          # KA = 17, KB = 3, two shared events, PT first
        _pair_row(n, ev[:18], [0, 19], 0),            # KA = 18, KB = 2, unknown order
        _pair_row(n, ev[13:18], ev[:17], 2),          # KA = 5, KB = 17 (wide MT side), MT first
    ])
    params = np.concatenate([th.ravel(), dp, dm])
    for r in range(rows.shape[0]):
        h = Handle(rows[r:r + 1])
        s, g = h.eval_weighted(params, 1.0, 1.0)
        h.close()
        out = ld.patient_value_grad(th, dp, dm, rows[r])
        assert abs(s - out[1]) <= TOL * abs(out[1]), r
        assert rel_err(g, np.concatenate([out[2].ravel(), out[3], out[4]])) <= TOL, r


# ---- the configurations the bench numbers are quoted on ------------------------------------------------------------------

def _restated_row(args):
    from oracle import reference_restated as rr
    th, dp, dm, row = args
    r = rr.patient_value_grad(th, dp, dm, row, want_grad=True)
    return None if r is None else (float(r[1]), np.concatenate([np.asarray(r[2], float).ravel(), np.asarray(r[3], float), np.asarray(r[4], float)]))


@pytest.mark.parametrize("n,n_dat", [(25, 100000), (20, 10000)])
def test_bench_dataset_rows_against_restated_reference(n, n_dat):
    """Rows of the VERY datasets bench.py times (SYN-v1, BASELINE configs[3] and configs[2]): a stratified sample with every
    (type, k) stratum up to 2^13 states (2^14 for pairs) against oracle/reference_restated.py -- per-row log-likelihood through
    mmh_per_patient, gradient sums through mmh_eval_weighted."""
    import multiprocessing as mp
    import os
    from metmhn_b200 import Handle
    from metmhn_b200.simulate import syn_v1
    d = syn_v1(n, n_dat, 1000 * n + 3)
    dat, ep = d["dat"], d["eval_point"]
    n_tot = n + 1
    th, dp, dm = ep[:n_tot * n_tot].reshape(n_tot, n_tot), ep[n_tot * n_tot:n_tot * (n_tot + 1)], ep[n_tot * (n_tot + 1):]
    typ = dat[:, -1]
    pt = dat[:, 0:2 * n:2].astype(int).sum(axis=1)
    mt = dat[:, 1:2 * n:2].astype(int).sum(axis=1)
    k = np.where(typ == 3, pt + mt + 1, np.where(typ == 2, mt + 1, pt + dat[:, 2 * n]))
    pick = []
    for t in range(4):
        for kk in range(0, 15 if t == 3 else 14):
            idx = np.nonzero((typ == t) & (k == kk))[0]
            pick.extend(idx[:2 if kk >= 12 else 3].tolist())
    pick = np.asarray(sorted(pick))
    assert len({(int(typ[i]), int(k[i])) for i in pick}) >= 45
    sub = np.ascontiguousarray(dat[pick])
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
        ref = pool.map(_restated_row, [(th, dp, dm, r) for r in sub], chunksize=1)
    h = Handle(sub)
    lp = h.per_patient(ep)
    s, g = h.eval_weighted(ep, 1.0, 1.0)
    S, G = 0.0, np.zeros(n_tot * (n_tot + 2))
    for i, r in enumerate(ref):
        assert r is not None
        assert abs(lp[i] - r[0]) <= TOL * abs(r[0]), (i, int(typ[pick[i]]), int(k[pick[i]]))
        S += r[0]
        G += r[1]
    assert abs(s - S) <= TOL * abs(S)
    assert rel_err(g, G) <= TOL


def test_tile_kernels_against_reference_golden_on_bench_rows():
    """tests/golden/golden_big.npz: rows of the bench datasets with 2^13 ... 2^18-state lattices, evaluated by the UNMODIFIED
    reference under the shim (tests/golden/make_golden_big.py).  Every row is solved by the big-tier kernels (tile solve,
    fused group-B statistics, product-form marginals), so this compares them with reference-derived numbers directly."""
    import os
    from metmhn_b200 import Handle
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_big.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/golden_big.npz has not been generated (tests/golden/make_golden_big.py)")
    gb = np.load(path)
    for c in [str(x) for x in gb["cases"]]:
        rows, ep = gb[f"{c}/rows"], gb[f"{c}/eval_point"]
        n_tot = (rows.shape[1] - 3) // 2 + 1
        bits = rows[:, :2 * (n_tot - 1) + 1].astype(int).sum(axis=1)
        assert bits.max() >= 16                         # every row runs on the big-tier kernels (2^13 ... 2^18 states)
        lp_all = Handle(rows).per_patient(ep)
        for r in range(rows.shape[0]):
            h = Handle(rows[r:r + 1])
            s, g = h.eval_weighted(ep, 1.0, 1.0)
            h.close()
            ref = gb[f"{c}/row_logp"][r]
            assert abs(s - ref) <= TOL * abs(ref), (c, r)
            assert abs(lp_all[r] - ref) <= TOL * abs(ref), (c, r)
            for got, key in zip(_split(g, n_tot), ("row_g", "row_gdp", "row_gdm")):
                want = gb[f"{c}/{key}"][r]
                if np.abs(want).max() == 0.0:
                    assert np.abs(got).max() == 0.0, (c, r, key)
                else:
                    assert rel_err(got, want) <= TOL, (c, r, key)


def test_per_patient_mirrors_of_the_reference_likelihood_module(golden):
    """Every function of metmhn_b200/likelihood.py (the mirrors of metmhn/jx/likelihood.py:286-730 and one_event.py) against the
    per-row goldens of the reference: each row of `kinds_n*` is sent through the function the reference's dispatch would call."""
    from metmhn_b200 import likelihood as lk
    called = set()
    for c in [str(x) for x in golden["cases"]]:
        if not c.startswith("kinds_n") or f"{c}/row_logp" not in golden.files:
            continue
        th, dp, dm, dat, _ = _case(golden, c)
        n = th.shape[0] - 1
        for r in range(dat.shape[0]):
            row = dat[r]
            typ, order = int(row[-1]), int(row[-2])
            ref_lp = golden[f"{c}/row_logp"][r]
            ref = (golden[f"{c}/row_g"][r], golden[f"{c}/row_gdp"][r], golden[f"{c}/row_gdm"][r])
            state_pt = np.append(row[0:2 * n:2], row[2 * n])
            state_mt = np.append(row[1:2 * n:2], row[2 * n])
            n_prim, n_met = int(state_pt.sum()), int(state_mt.sum())
            if typ in (0, 1):
                if typ == 0 and n_prim == 0:
                    lp, g, gdp = lk._grad_prim_obs_az(th)
                    lp2 = lk._lp_prim_obs_az(th)
                    called.update(["_grad_prim_obs_az", "_lp_prim_obs_az"])
                else:
                    lp, g, gdp = lk._grad_prim_obs(th, dp, state_pt, n_prim)
                    lp2 = lk._lp_prim_obs(th, dp, state_pt, n_prim)
                    called.update(["_grad_prim_obs", "_lp_prim_obs"])
                gdm = np.zeros(n + 1)
            elif typ == 2:
                lp, g, gdp, gdm = lk._grad_met_obs(th, dp, dm, state_mt, n_met)
                lp2 = lk._lp_met_obs(th, dp, dm, state_mt, n_met)
                called.update(["_grad_met_obs", "_lp_met_obs"])
            else:
                o = order if order in (0, 1) else 2
                fg = (lk._g_coupled_0, lk._g_coupled_1, lk._g_coupled_2)[o]
                fl = (lk._lp_coupled_0, lk._lp_coupled_1, lk._lp_coupled_2)[o]
                lp, g, gdp, gdm = fg(th, dp, dm, row[:2 * n + 1], n_prim, n_met)
                lp2 = fl(th, dp, dm, row[:2 * n + 1], n_prim, n_met)
                called.update([fg.__name__, fl.__name__])
            assert abs(lp - ref_lp) <= TOL * abs(ref_lp) and abs(lp2 - ref_lp) <= TOL * abs(ref_lp), (c, r)
            for got, want in zip((g, gdp, gdm), ref):
                if np.abs(want).max() == 0.0:
                    assert np.abs(got).max() == 0.0, (c, r)
                else:
                    assert rel_err(got, want) <= TOL, (c, r)
    assert called == {"_g_coupled_0", "_g_coupled_1", "_g_coupled_2", "_lp_coupled_0", "_lp_coupled_1", "_lp_coupled_2",
                      "_grad_prim_obs", "_lp_prim_obs", "_grad_prim_obs_az", "_lp_prim_obs_az", "_grad_met_obs", "_lp_met_obs"}
