"""Config C1 of BASELINE.json: the reference's own runnable workload (examples/analysis.py on data/luad, all 28 events,
fixed lambda, no CV) on the GPU path.  Uses the fixture tests/golden/luad_dat.npz (derived from the reference CSVs)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metmhn_b200 as mm
from metmhn_b200.utility import indep

dat = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "luad_dat.npz"))["dat"]
th0, dp0, dm0 = indep(dat)
h = mm.regularized_optimization.dataset_handle(dat)
x0 = np.concatenate([th0.ravel(), dp0, dm0])
h.value_grad(x0, 0.65)
t = time.perf_counter()
for _ in range(5):
    h.value_grad(x0, 0.65)
t_eval = (time.perf_counter() - t) / 5
n_eval = [0]
def penal(p, n):
    n_eval[0] += 1
    return mm.symmetric_penal(p, n)
t = time.perf_counter()
th, dp, dm = mm.learn_mhn(th0, dp0, dm0, dat, 0.65, penal, 1e-3, opt_ftol=1e-5, opt_v=False)
t_fit = time.perf_counter() - t
x = np.concatenate([th.ravel(), dp, dm])
f, g = mm.score_and_grad_reg(x, dat, 0.65, mm.symmetric_penal, 1e-3)
print(json.dumps({"workload": "LUAD, 4852 patients, 28 events, perc_met 0.65, lambda 1e-3, ftol 1e-5",
                  "value_grad_ms": 1e3 * t_eval, "patients_per_s": dat.shape[0] / t_eval, "fit_s": t_fit,
                  "evaluations": n_eval[0], "objective": float(f), "max_abs_grad": float(np.abs(g).max()),
                  "stats": {k: v for k, v in h.stats().items() if k in ("n_spaces", "n_chunks", "states_value_grad", "n_launches")}}))
