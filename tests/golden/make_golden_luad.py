"""Golden vectors on the reference's real workload (LUAD, 28 events): `indep`, `score`, `score_and_grad` of the
UNMODIFIED reference (on the NumPy JAX shim) for a stratified subset of rows small enough for the reference
algorithm.  Run in the build container:  python tests/golden/make_golden_luad.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_under_shim  # noqa: E402

regopt, lik, one, van, kv = ref_under_shim.load()
import importlib
import jax.numpy as jnp
utils = importlib.import_module("metmhn.Utilityfunctions")

dat = np.load(os.path.join(HERE, "luad_dat.npz"))["dat"]
n = (dat.shape[1] - 3) // 2
th0, dp0, dm0 = utils.indep(jnp.array(dat))
rng = np.random.default_rng(2024)
th = np.asarray(th0) + rng.normal(0, 0.15, (n + 1, n + 1))
dp = np.asarray(dp0) + rng.normal(0, 0.2, n + 1)
dm = np.asarray(dm0) + rng.normal(0, 0.2, n + 1)
typ = dat[:, -1]
bits = np.where(typ == 3, dat[:, :2 * n + 1].astype(int).sum(axis=1),
                np.where(typ == 2, dat[:, 1:2 * n:2].astype(int).sum(axis=1) + 1, dat[:, 0:2 * n + 1:2].astype(int).sum(axis=1)))
pick = []
for t, kmax, cnt in ((0, 8, 6), (1, 8, 8), (2, 8, 8), (3, 10, 14)):
    idx = np.nonzero((typ == t) & (bits <= kmax))[0]
    idx = idx[np.argsort(-bits[idx], kind="stable")][:cnt // 2].tolist() + idx[:cnt - cnt // 2].tolist()
    pick += idx
sub = np.ascontiguousarray(dat[sorted(set(pick))])
s, g, a, b = regopt.score_and_grad(jnp.array(th), jnp.array(dp), jnp.array(dm), jnp.array(sub), 0.65)
s2 = regopt.score(jnp.array(th), jnp.array(dp), jnp.array(dm), jnp.array(sub), 0.65)
np.savez_compressed(os.path.join(HERE, "golden_luad.npz"), indep_theta=np.asarray(th0), theta=th, d_p=dp, d_m=dm, rows=sub,
                    score=np.float64(np.asarray(s).reshape(-1)[0]), g=np.asarray(g), gdp=np.asarray(a), gdm=np.asarray(b),
                    score_only=np.float64(np.asarray(s2).reshape(-1)[0]))
print("rows", sub.shape, "score", s, s2)
