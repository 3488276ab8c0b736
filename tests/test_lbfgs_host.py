"""The in-library optimiser (metmhn_b200/csrc/mmh_lbfgs.hpp, the loop behind `learn_mhn(..., optimizer="native")`) against
SciPy's L-BFGS-B -- the driver the reference uses (metmhn/regularized_optimization.py:328) -- on the same problems: same
optimum, comparable number of iterations and evaluations.  Compiled with g++, no GPU needed."""
import json
import os
import subprocess

import numpy as np
import scipy.optimize as opt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _native(tmp_path):
    exe = str(tmp_path / "lbfgs_host_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "lbfgs_host_test.cpp")], check=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, check=True).stdout
    return {d["name"]: d for d in map(json.loads, out.strip().splitlines())}


def test_native_lbfgs_reaches_the_same_optima_as_scipy(tmp_path):
    got = _native(tmp_path)

    def rosen(x):
        a, b = x[1:] - x[:-1] ** 2, 1.0 - x[:-1]
        g = np.zeros_like(x)
        g[:-1] += -400.0 * a * x[:-1] - 2.0 * b
        g[1:] += 200.0 * a
        return float(np.sum(100.0 * a * a + b * b)), g
    x0 = np.where(np.arange(50) % 2 == 1, 1.0, -1.2)
    ref = opt.minimize(rosen, x0, jac=True, method="L-BFGS-B", options={"maxiter": 100000, "ftol": 1e-12, "gtol": 1e-8})
    r = got["rosenbrock50"]
    assert r["status"] in (0, 1) and r["f"] <= 1e-10 and abs(r["x0"] - 1.0) <= 1e-5 and ref.fun <= 1e-8
    assert r["ev"] <= 2 * ref.nfev + 20

    n = 899
    i = np.arange(n)
    a, w = np.sin(0.37 * i) + 0.1 * (i % 7), 1.0 + 499.0 * i / (n - 1)
    ref = opt.minimize(lambda x: (0.5 * float(np.sum(w * (x - a) ** 2)), w * (x - a)), np.zeros(n), jac=True, method="L-BFGS-B",
                       options={"maxiter": 100000, "ftol": 1e-14, "gtol": 1e-10})
    r = got["quad899"]
    assert r["status"] in (0, 1) and r["f"] <= 1e-12 and abs(r["x0"] - a[0]) <= 1e-6
    assert r["it"] <= 1.5 * ref.nit + 10 and r["ev"] <= 1.5 * ref.nfev + 10

    C = np.array([[np.sin(1.3 * rr + 0.7 * ii) + (0.5 if (rr + ii) % 3 == 0 else -0.25) for ii in range(40)] for rr in range(120)])

    def logi(x):
        z = C @ x
        return float(np.sum(np.logaddexp(0.0, z)) + 0.05 * x @ x), C.T @ (1.0 / (1.0 + np.exp(-z))) + 0.1 * x
    tight = opt.minimize(logi, np.zeros(40), jac=True, method="L-BFGS-B", options={"maxiter": 100000, "ftol": 1e-13, "gtol": 1e-9})
    loose = opt.minimize(logi, np.zeros(40), jac=True, method="L-BFGS-B", options={"maxiter": 100000, "ftol": 1e-4})
    r = got["logistic40_tight"]
    assert abs(r["f"] - tight.fun) <= 1e-9 * abs(tight.fun) and r["ev"] <= 1.5 * tight.nfev + 10
    r = got["logistic40_ftol1e-4"]        # learn_mhn's default tolerance: both stop within ftol of the optimum
    assert r["status"] == 0 and abs(r["f"] - tight.fun) <= 5e-4 * abs(tight.fun) and abs(loose.fun - tight.fun) <= 5e-4 * abs(tight.fun)
    assert abs(r["it"] - loose.nit) <= max(3, loose.nit // 2)
