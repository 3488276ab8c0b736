"""Config C1 of BASELINE.json: the reference's own runnable workload (examples/analysis.py on data/luad, all 28 events,
fixed lambda, no CV) on the GPU path, with both optimiser loops (in-library L-BFGS and SciPy's), and the model CSV the
reference writes (analysis.py:115-120).  Uses the fixture tests/golden/luad_dat.npz (derived from the reference CSVs)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metmhn_b200 as mm
from metmhn_b200 import regularized_optimization as ro
from metmhn_b200.utility import indep, write_model_csv, read_model_csv

dat = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "luad_dat.npz"))["dat"]
th0, dp0, dm0 = indep(dat)
h = ro.dataset_handle(dat)
x0 = np.concatenate([th0.ravel(), dp0, dm0])
h.value_grad(x0, 0.65)
t = time.perf_counter()
for _ in range(5):
    h.value_grad(x0, 0.65)
t_eval = (time.perf_counter() - t) / 5
out = {"workload": "LUAD, 4852 patients, 28 events, perc_met 0.65, lambda 1e-3, ftol 1e-5",
       "value_grad_ms": 1e3 * t_eval, "patients_per_s": dat.shape[0] / t_eval}
fits = {}
for name in ("native", "scipy"):
    mm.learn_mhn(th0, dp0, dm0, dat, 0.65, mm.symmetric_penal, 1e-3, opt_ftol=1e-5, opt_v=False, optimizer=name)     # warm-up
    t = time.perf_counter()
    th, dp, dm = mm.learn_mhn(th0, dp0, dm0, dat, 0.65, mm.symmetric_penal, 1e-3, opt_ftol=1e-5, opt_v=False, optimizer=name)
    t_fit = time.perf_counter() - t
    x = np.concatenate([th.ravel(), dp, dm])
    f, g = mm.score_and_grad_reg(x, dat, 0.65, mm.symmetric_penal, 1e-3)
    info = dict(ro.LAST_FIT)
    fits[name] = {"fit_s": t_fit, "iterations": info["iterations"], "evaluations": info["evaluations"],
                  "ms_per_iteration": 1e3 * t_fit / max(info["iterations"], 1), "ms_per_evaluation": 1e3 * t_fit / max(info["evaluations"], 1),
                  "objective": float(f), "max_abs_grad": float(np.abs(g).max())}
    if name == "native":
        names = [f"E{i}" for i in range(th.shape[0] - 1)] + ["Seeding"]
        path = os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "luad_model.csv")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        write_model_csv(path, th, dp, dm, names)
        th2, dp2, dm2, _ = read_model_csv(path)
        out["model_csv"] = {"path": "gpurun_out/luad_model.csv", "rows": int(2 + th.shape[0]), "round_trip_exact": bool(np.array_equal(th2, th) and np.array_equal(dp2, dp))}
out["fits"] = fits
out["objective_native_minus_scipy"] = fits["native"]["objective"] - fits["scipy"]["objective"]
# the same optimum at a tight tolerance
xs = {}
for name in ("native", "scipy"):
    th, dp, dm = mm.learn_mhn(th0, dp0, dm0, dat, 0.65, mm.symmetric_penal, 1e-3, opt_ftol=1e-13, opt_v=False, optimizer=name)
    xs[name] = float(mm.score_and_grad_reg(np.concatenate([th.ravel(), dp, dm]), dat, 0.65, mm.symmetric_penal, 1e-3)[0])
out["objective_at_ftol_1e-13"] = xs
out["stats"] = {k: v for k, v in h.stats().items() if k in ("n_spaces", "n_chunks", "states_value_grad", "n_launches")}
print(json.dumps(out))
