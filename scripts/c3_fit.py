"""Config C3 of BASELINE.json: synthetic n = 20 events, 10 000 patients (coupled PT/MT, primary-only, met-only), a FULL
L-BFGS fit on one B200: `learn_mhn(symmetric_penal, w_penal = 1e-3, opt_ftol = 1e-5)` from the independence model, with
the in-library optimiser (default) and with SciPy's driver.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metmhn_b200 as mm
from metmhn_b200 import regularized_optimization as ro
from metmhn_b200.simulate import syn_v1
from metmhn_b200.utility import indep

n, n_dat = 20, 10000
d = syn_v1(n, n_dat, 1000 * n + 3)
dat = d["dat"]
t = time.perf_counter()
h = ro.dataset_handle(dat)
t_create = time.perf_counter() - t
th0, dp0, dm0 = indep(dat)
x0 = np.concatenate([th0.ravel(), dp0, dm0])
f0 = float(mm.score_and_grad_reg(x0, h, 0.65, mm.symmetric_penal, 1e-3)[0])
out = {"config": "C3: SYN-v1 n=20, 10000 patients, learn_mhn(symmetric_penal, w_penal=1e-3, opt_ftol=1e-5), start = indep(dat), 1 B200",
       "dataset_upload_and_plan_s": t_create, "objective_start": f0, "fits": {}}
for name in ("native", "scipy"):
    t = time.perf_counter()
    th, dp, dm = mm.learn_mhn(th0, dp0, dm0, h, 0.65, mm.symmetric_penal, 1e-3, opt_ftol=1e-5, opt_v=False, optimizer=name)
    t_fit = time.perf_counter() - t
    info = dict(ro.LAST_FIT)
    x = np.concatenate([th.ravel(), dp, dm])
    f, g = mm.score_and_grad_reg(x, h, 0.65, mm.symmetric_penal, 1e-3)
    out["fits"][name] = {"fit_s": t_fit, "iterations": info["iterations"], "evaluations": info["evaluations"],
                         "ms_per_evaluation": 1e3 * t_fit / max(info["evaluations"], 1), "objective": float(f),
                         "max_abs_grad": float(np.abs(g).max()),
                         "patients_per_s_during_fit": n_dat * info["evaluations"] / t_fit}
out["stats"] = {k: v for k, v in h.stats().items() if k in ("n_spaces", "n_chunks", "states_value_grad", "n_launches", "last_ms")}
print(json.dumps(out))
