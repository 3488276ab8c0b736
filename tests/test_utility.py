"""CPU tests of the callers of the hot path: indep, ingestion fixture, oracle on the LUAD golden subset."""
import os

import numpy as np

from conftest import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _luad():
    return np.load(os.path.join(ROOT, "tests", "golden", "luad_dat.npz"))["dat"]


def test_luad_fixture_matches_survey_counts():
    dat = _luad()
    assert dat.shape == (4852, 59) and dat.dtype == np.int8
    typ, cnt = np.unique(dat[:, -1], return_counts=True)
    assert dict(zip(typ.tolist(), cnt.tolist())) == {0: 595, 1: 1677, 2: 2127, 3: 453}
    orders = dict(zip(*[x.tolist() for x in np.unique(dat[dat[:, -1] == 3, -2], return_counts=True)]))
    assert orders == {-99: 1, 0: 113, 1: 259, 2: 80}


def test_indep_matches_reference():
    from metmhn_b200.utility import indep
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_luad.npz"))
    th, dp, dm = indep(_luad())
    assert np.abs(th - g["indep_theta"]).max() < 1e-12 and not dp.any() and not dm.any()


def test_oracle_on_luad_subset_matches_reference():
    from oracle import lattice_direct as ld
    from oracle import reference_restated as rr
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_luad.npz"))
    for mod in (rr, ld):
        s, gg, a, b = mod.score_and_grad(g["theta"], g["d_p"], g["d_m"], g["rows"], 0.65)
        assert abs(s - g["score"]) <= 1e-10 * abs(g["score"])
        assert rel_err(gg, g["g"]) <= 1e-10 and rel_err(a, g["gdp"]) <= 1e-10 and rel_err(b, g["gdm"]) <= 1e-10


def test_model_csv_round_trip_and_reference_layout(tmp_path):
    """`row_stack(d_p, d_m, theta)` with the event names as columns (examples/analysis.py:115-120)."""
    import os
    from metmhn_b200.utility import read_model_csv, write_model_csv
    rng = np.random.default_rng(3)
    n_tot = 5
    th, dp, dm = rng.normal(size=(n_tot, n_tot)), rng.normal(size=n_tot), rng.normal(size=n_tot)
    names = ["TP53 (M)", "KRAS (M)", "EGFR (M)", "STK11 (M)", "Seeding"]
    p = str(tmp_path / "model.csv")
    write_model_csv(p, th, dp, dm, names)
    lines = open(p).read().splitlines()
    assert lines[0] == "," + ",".join(names) and len(lines) == 1 + 2 + n_tot and lines[1].startswith("0,") and lines[3].startswith("2,")
    th2, dp2, dm2, names2 = read_model_csv(p)
    assert names2 == names and np.array_equal(th2, th) and np.array_equal(dp2, dp) and np.array_equal(dm2, dm)
    ref = "/root/reference/results/luad/luad_g14_20muts.csv"
    if os.path.exists(ref):                       # the published model has the same layout (23 rows = d_p, d_m, 21 x 21 theta)
        th3, dp3, dm3, names3 = read_model_csv(ref)
        assert th3.shape == (21, 21) and dp3.shape == (21,) and names3[-1] == "Seeding"
