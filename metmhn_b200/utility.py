"""Callers of the hot path ("next" rows of SURVEY.md 8f): the independence-model initialiser, the k-fold x lambda
cross-validation sweep and the CSV ingestion of the reference's own runnable workload.

Mirrors `metmhn/Utilityfunctions.py:157-231` (`indep`, `cross_val`) and `examples/analysis.py:49-72` (ingestion); the
likelihood evaluations run on the GPU through `metmhn_b200.regularized_optimization`.
"""
from __future__ import annotations

import logging

import numpy as np

from .regularized_optimization import learn_mhn, score


def indep(dat):
    """Initial estimate: log-odds of every event on the diagonal, zero interactions and diagnosis effects
    (`Utilityfunctions.py:157-183`)."""
    dat = np.asarray(dat)
    n = (dat.shape[1] - 3) // 2
    n_coupled = int((dat[:, -1] == 3).sum())
    n_single = dat.shape[0] - n_coupled
    theta = np.zeros((n + 1, n + 1))
    for i in range(n):
        count = float(dat[:, 2 * i].astype(np.int64).sum() + dat[:, 2 * i + 1].astype(np.int64).sum())
        theta[i, i] = -1e10 if count == 0 else np.log(count / (2 * n_coupled + n_single - count + 1e-10))
    seeded = float(dat[:, -3].astype(np.int64).sum())
    theta[n, n] = np.log(seeded / (n_coupled + n_single - seeded + 1e-10))
    return theta, np.zeros(n + 1), np.zeros(n + 1)


def cross_val(dat, penal_fun, splits, n_folds: int, m_p_corr: float, seed: int = 42):
    """k-fold cross-validation over the penalty weights `splits` (`Utilityfunctions.py:186-231`).
    Returns an (n_folds, len(splits)) array of held-out scores.  The shuffle uses NumPy's PCG64 (the reference uses
    jax.random, whose stream cannot be reproduced without JAX); everything else follows the reference loop."""
    dat = np.asarray(dat)
    splits = np.asarray(splits, dtype=float)
    shuffled = dat[np.random.Generator(np.random.PCG64(seed)).permutation(dat.shape[0])]
    runs = np.zeros((n_folds, splits.shape[0]))
    batch = int(np.ceil(dat.shape[0] / n_folds))
    for i, lam in enumerate(splits):
        for fold in range(n_folds):
            start, stop = batch * fold, min(batch * (fold + 1), dat.shape[0])
            train = np.ascontiguousarray(np.concatenate([shuffled[:start], shuffled[stop:]]))
            test = np.ascontiguousarray(shuffled[start:stop])
            th0, dp0, dm0 = indep(train)
            th, dp, dm = learn_mhn(th0, dp0, dm0, train, m_p_corr, penal_fun, lam, opt_v=False)
            runs[fold, i] = score(th, dp, dm, test, m_p_corr)
            logging.info("Lambda: %s Fold: %d Test Score: %s", lam, fold, runs[fold, i])
    return runs


def cross_val_distributed(dat, penal_fun, splits, n_folds: int, m_p_corr: float, seed: int = 42,
                          rank: int = 0, world: int = 1, device: int = 0, group=None, fit_and_score=None):
    """The same sweep with the (lambda, fold) jobs dealt round-robin to the ranks of a torch.distributed group (one
    process per GPU, BASELINE config 5): the 25 fits of the default 5 x 5 sweep are independent, so there is no
    data-path collective, only one all-reduce of the (n_folds, len(splits)) result table at the end.
    `fit_and_score(train, test, lam) -> float` defaults to `learn_mhn` + `score` on this rank's GPU."""
    dat = np.asarray(dat)
    splits = np.asarray(splits, dtype=float)
    shuffled = dat[np.random.Generator(np.random.PCG64(seed)).permutation(dat.shape[0])]
    batch = int(np.ceil(dat.shape[0] / n_folds))
    if fit_and_score is None:
        from .regularized_optimization import dataset_handle

        def fit_and_score(train, test, lam):
            th0, dp0, dm0 = indep(train)
            htrain = dataset_handle(train, device=device)
            th, dp, dm = learn_mhn(th0, dp0, dm0, htrain, m_p_corr, penal_fun, lam, opt_v=False)
            return dataset_handle(test, device=device).value(np.concatenate([th.ravel(), dp, dm]), m_p_corr)

    runs = np.zeros((n_folds, splits.shape[0]))
    jobs = [(i, f) for i in range(splits.shape[0]) for f in range(n_folds)]
    for j, (i, fold) in enumerate(jobs):
        if j % world != rank:
            continue
        start, stop = batch * fold, min(batch * (fold + 1), dat.shape[0])
        train = np.ascontiguousarray(np.concatenate([shuffled[:start], shuffled[stop:]]))
        test = np.ascontiguousarray(shuffled[start:stop])
        runs[fold, i] = fit_and_score(train, test, float(splits[i]))
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(runs)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        runs = t.cpu().numpy()
    return runs


def categorize(paired, meta_status):
    """Row type from the annotation columns (`Utilityfunctions.py:98-113`); None = unusable row."""
    if paired == 0:
        return {"absent": 0, "present": 1, "isMetastasis": 2}.get(meta_status, None if meta_status == "unknown" else -1)
    if paired == 1:
        return 3
    return None


def read_events_csv(events_csv: str, annot_csv: str):
    """Build the int8 data matrix from the two CSVs of `data/luad` exactly as `examples/analysis.py:49-72` does.
    Returns (dat int8 (n, 2n+3), event names)."""
    import pandas as pd
    annot = pd.read_csv(annot_csv)
    mut = pd.read_csv(events_csv).rename(columns={"Unnamed: 0": "patientID"})
    df = pd.merge(mut, annot.loc[:, ["patientID", "metaStatus"]], on="patientID")
    muts = list(df.columns[1:-4])
    typ = [categorize(p, m) for p, m in zip(df["paired"], df["metaStatus"])]
    age_m = pd.to_numeric(df["M.AgeAtSeqRep"], errors="coerce").to_numpy()
    age_p = pd.to_numeric(df["P.AgeAtSeqRep"], errors="coerce").to_numpy()
    diff = age_m - age_p
    order = np.where(np.isnan(diff), -99, np.where(diff < 0, 2, np.where(diff > 0, 1, 0)))
    keep = np.array([t is not None for t in typ])
    geno = df.loc[keep, muts].to_numpy(dtype=np.int8)
    t = np.array([x for x in typ if x is not None], dtype=np.int8)
    seeding = np.where(t == 0, 0, 1).astype(np.int8)
    dat = np.concatenate([geno, seeding[:, None], order[keep].astype(np.int8)[:, None], t[:, None]], axis=1)
    names = [c.split(".", 1)[1] for c in muts[::2]] + ["Seeding"]
    return np.ascontiguousarray(dat.astype(np.int8)), names


def write_model_csv(path: str, theta, d_p, d_m, events):
    """Write a fitted model the way `examples/analysis.py:115-120` does: a pandas CSV whose columns are the event names
    (seeding last) and whose rows are `row_stack(d_p, d_m, theta)`, indexed 0 ... n+2 (the layout of
    `results/luad/luad_g14_20muts.csv`)."""
    import pandas as pd
    theta = np.asarray(theta, dtype=np.float64)
    n_tot = theta.shape[0]
    if theta.shape != (n_tot, n_tot) or len(events) != n_tot:
        raise ValueError("theta must be (n+1, n+1) and `events` must name its n+1 columns (seeding last)")
    table = np.vstack([np.asarray(d_p, dtype=np.float64).reshape(1, -1), np.asarray(d_m, dtype=np.float64).reshape(1, -1), theta])
    pd.DataFrame(table, columns=list(events)).to_csv(path)


def read_model_csv(path: str):
    """Inverse of `write_model_csv`; also reads the reference's published models (`results/luad/*.csv`).
    Returns (theta, d_p, d_m, events)."""
    import pandas as pd
    df = pd.read_csv(path, index_col=0, float_precision="round_trip")
    table = df.to_numpy(dtype=np.float64)
    return table[2:].copy(), table[0].copy(), table[1].copy(), list(df.columns)
