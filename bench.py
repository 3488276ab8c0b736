#!/usr/bin/env python
"""Benchmark of the hot path: value_grad (score_and_grad) patients/sec, FP64.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU

Workload (BASELINE.json configs[3], SYN-v1 of SURVEY.md 8d): n = 25 events, 100 000 synthetic patients
(mixed paired / unpaired), one step = one likelihood + gradient evaluation of the WHOLE dataset at the
evaluation point P3.  With N > 1 (torchrun, one rank per GPU) the patients are sharded over the ranks by a
cost model and the step ends with one NCCL all-reduce of 1 + (n+1)(n+3) doubles: strong scaling.
The dataset is uploaded once when the handle is created (it is constant over the L-BFGS iterations of
`learn_mhn`, like model weights); the per-step input is the parameter vector, the per-step output the
score and gradient.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "value_grad patients/sec (FP64)"
PERC_MET = 0.65


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def base_config(args):
    return {"workload": f"SYN-v1 (SURVEY 8d) n={args.n} events, {args.patients} patients "
                        f"(11.5% type 0, then 10.7% paired / 38.6% PT-only / rest MT-only), value_grad at P3",
            "n_events": args.n, "n_patients": args.patients, "perc_met": PERC_MET}


def workload(args):
    from metmhn_b200.simulate import syn_v1
    d = syn_v1(args.n, args.patients, 1000 * args.n + 3)
    return d


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---- CPU baseline (reference algorithm, NumPy restatement) ------------------------------------------------
def _cpu_one(args):
    """One row through the reference algorithm; returns its time and its result (kept for the parity check)."""
    from oracle import reference_restated as rr
    i, th, dp, dm, row = args
    t = time.perf_counter()
    r = rr.patient_value_grad(th, dp, dm, row, want_grad=True)
    dt = time.perf_counter() - t
    if r is None:
        return i, dt, 0.0, None
    is0, lp, g, gdp, gdm = r
    return i, dt, float(lp), np.concatenate([np.asarray(g, dtype=float).ravel(), np.asarray(gdp, dtype=float).ravel(),
                                             np.asarray(gdm, dtype=float).ravel()])


def row_strata(dat):
    """(type, k) of every row and the reference's work estimate W = sum over the row's spaces of
    N_eff * k_eff^2 * (k_eff + 1) ((k+1) Jacobi sweeps of ~k^2 passes over N_eff states)."""
    n = (dat.shape[1] - 3) // 2
    typ = dat[:, -1].astype(np.int64)
    pt = dat[:, 0:2 * n:2].astype(np.int64).sum(axis=1)
    mt = dat[:, 1:2 * n:2].astype(np.int64).sum(axis=1)
    seed = dat[:, 2 * n].astype(np.int64)

    def w(k):
        k = np.maximum(k, 0).astype(np.float64)
        return np.exp2(k) * k * k * (k + 1.0)
    k = np.where(typ == 3, pt + mt + 1, np.where(typ == 2, mt + 1, pt + seed))
    W = np.where(typ == 3, w(pt + mt) + w(mt) + w(pt), w(k))          # joint + both second-phase spaces (upper bound for order 1/2)
    W = np.where((typ < 0) | (typ > 3), 0.0, W)
    return typ, k, W


def row_passes(dat):
    """P = sum over the row's spaces of k_eff^2 (k_eff + 1): the number of vector passes behind W (W = N_eff * P per space)."""
    n = (dat.shape[1] - 3) // 2
    typ = dat[:, -1].astype(np.int64)
    pt = dat[:, 0:2 * n:2].astype(np.int64).sum(axis=1).astype(np.float64)
    mt = dat[:, 1:2 * n:2].astype(np.int64).sum(axis=1).astype(np.float64)
    seed = dat[:, 2 * n].astype(np.float64)
    p = lambda k: k * k * (k + 1.0)
    P = np.where(typ == 3, p(pt + mt) + p(mt) + p(pt), np.where(typ == 2, p(mt + 1), p(pt + seed)))
    return np.where((typ < 0) | (typ > 3), 0.0, P)


def _cost_model(timed):
    """cpu seconds of a row ~ c0 + c_pass * P + c_elem * W (P = vector passes, W = passes x elements): non-negative
    least squares on the per-stratum means, relative residuals.  The per-pass term keeps the interpreter overhead of
    the small lattices out of the per-element cost that the extrapolation to the big ones rests on."""
    from scipy.optimize import nnls
    vals = [v for v in timed.values() if v[0] > 0 and v[1] > 0]
    t = np.array([v[1] / v[0] for v in vals])
    A = np.array([[1.0, v[3] / v[0], v[2] / v[0]] for v in vals])
    coef, _ = nnls(A / t[:, None], np.ones(len(vals)))
    return float(coef[0]), float(coef[1]), float(coef[2])


def cpu_baseline(d, budget_s, per_stratum=24, seed=0):
    """Time the reference algorithm (oracle/reference_restated.py: per-event Kronecker shuffles, (k+1) Jacobi sweeps)
    on all host cores over a STRATIFIED sample of the workload: rows grouped by (type, k), cheapest strata first,
    up to `per_stratum` rows each, until the time budget is spent; the strata that were not reached (the 2^k tail,
    hours of NumPy per row) are extrapolated with the survey's work estimate W at the cost per unit of W measured on
    the two most expensive strata that were timed.  Returns the extrapolated whole-workload throughput and the
    per-row results of the sample (for the parity check against the GPU)."""
    import multiprocessing as mp
    dat = d["dat"]
    n = (dat.shape[1] - 3) // 2
    ep = d["eval_point"]
    n_tot = n + 1
    th, dp, dm = ep[:n_tot * n_tot].reshape(n_tot, n_tot), ep[n_tot * n_tot:n_tot * (n_tot + 1)], ep[n_tot * (n_tot + 1):]
    typ, k, W = row_strata(dat)
    P = row_passes(dat)
    rng = np.random.default_rng(seed)
    strata = {}
    for t in range(4):
        for kk in np.unique(k[typ == t]):
            rows = np.nonzero((typ == t) & (k == kk))[0]
            strata[(t, int(kk))] = rows
    order = sorted(strata, key=lambda key: float(W[strata[key][0]]))
    cores = os.cpu_count() or 1
    t_start = time.perf_counter()
    timed = {}                                  # stratum -> (rows timed, cpu seconds)
    res_idx, res_lp, res_g = [], [], []
    with mp.get_context("fork").Pool(cores) as pool:
        for key in order:
            if time.perf_counter() - t_start > budget_s:
                break
            rows = strata[key]
            take = rows if rows.shape[0] <= per_stratum else rng.choice(rows, per_stratum, replace=False)
            # expensive strata: only as many rows as the remaining budget allows (at least two, one per core at most)
            if timed:
                c0, cp, unit = _cost_model(timed)
                est = c0 + cp * float(P[take[0]]) + unit * float(W[take[0]])
                left = max(budget_s - (time.perf_counter() - t_start), 0.0)
                fit = int(left * cores / max(est, 1e-9))
                if fit < 2:
                    break
                take = take[:max(2, min(take.shape[0], fit, 2 * cores))]
            cpu = 0.0
            for i, dt, lp, g in pool.imap_unordered(_cpu_one, ((int(i), th, dp, dm, dat[i]) for i in take), chunksize=1):
                cpu += dt
                if g is not None:
                    res_idx.append(i); res_lp.append(lp); res_g.append(g)
            timed[key] = (int(take.shape[0]), cpu, float(W[take].sum()), float(P[take].sum()))
        pool.terminate()
    wall = time.perf_counter() - t_start
    # extrapolation: measured strata scale by their row count; the others by W at the cost per unit of W of the two
    # most expensive strata that were timed
    c0, cp, unit = _cost_model(timed)
    cpu_total, cpu_measured_part, rows_measured = 0.0, 0.0, 0
    for key, rows in strata.items():
        if key in timed:
            cnt, cpu = timed[key][:2]
            cpu_total += cpu / cnt * rows.shape[0]
            cpu_measured_part += cpu / cnt * rows.shape[0]
            rows_measured += rows.shape[0]
        else:
            cpu_total += c0 * rows.shape[0] + cp * float(P[rows].sum()) + unit * float(W[rows].sum())
    n_rows = int(sum(v[0] for v in timed.values()))
    est_wall = cpu_total / cores
    kmax = {t: max((kk for (tt, kk) in timed if tt == t), default=-1) for t in range(4)}
    return {"value": dat.shape[0] / est_wall, "unit": "patients/s", "cores": cores, "kind": "port",
            "extrapolated": True,
            "sample": f"stratified by (type, k): {n_rows} rows from {len(timed)} of {len(strata)} strata (every stratum up to "
                      f"k = {kmax} per type), {wall:.1f} s wall on {cores} cores; reference algorithm = "
                      f"oracle/reference_restated.py (NumPy restatement: JAX is not installable here), multiprocessing over "
                      f"rows; whole-workload time = measured strata scaled by their row counts ({rows_measured} rows, "
                      f"{cpu_measured_part / cores:.1f} s) + the remaining strata extrapolated with "
                      f"cpu-s per row = {c0:.2e} + {cp:.2e} * P + {unit:.2e} * W (non-negative least squares on the stratum means; "
                      f"P = sum k_eff^2 (k_eff+1) passes, W = sum N_eff k_eff^2 (k_eff+1)); total {est_wall:.0f} s",
            "rows": n_rows, "seconds": wall, "extrapolated_seconds_full_workload": est_wall,
            "measured_share_of_rows": rows_measured / max(dat.shape[0], 1),
            "row_index": np.asarray(res_idx, dtype=np.int64), "row_logp": np.asarray(res_lp),
            "row_grad_sum": np.sum(res_g, axis=0) if res_g else None}


def run_reference(args, rank, world):
    """Reference arm: the reference algorithm on the host cores, same workload, same metric.  A step is one bounded
    stratified sample (see cpu_baseline); the value is the whole-workload throughput extrapolated from it."""
    if rank != 0:
        return
    d = workload(args)
    for _ in range(args.warmup):
        cpu_baseline(d, 1.0, per_stratum=4)
    t, last, vals = [], None, []
    for s in range(args.steps):
        last = cpu_baseline(d, args.ref_step_seconds, seed=s)
        t.append(last["seconds"])
        vals.append(last["value"])
    val = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "patients/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / len(t),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(base_config(args), cpu_step="one bounded stratified sample of the workload's rows per step; value = "
                                                      "whole-workload patients/s extrapolated from it (cpu_baseline.sample)"),
            "cpu_baseline": {"value": val, "unit": "patients/s", "cores": last["cores"], "kind": "port",
                             "extrapolated": True, "sample": last["sample"]},
            "e2e": {"value": val, "unit": "patients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def parity_vs_oracle(cb, dat, ep, device):
    """The oracle's per-row results of the CPU sample against the CUDA path on the same rows of the bench dataset."""
    from metmhn_b200 import Handle
    rows = cb["row_index"]
    if rows.shape[0] == 0:
        return None
    h = Handle(np.ascontiguousarray(dat[rows]), device=device)
    lp = h.per_patient(ep)
    s, g = h.eval_weighted(ep, 1.0, 1.0, want_grad=True)
    t0 = time.perf_counter()
    for _ in range(3):
        h.eval_weighted(ep, 1.0, 1.0, want_grad=True)
    gpu_rate = rows.shape[0] * 3 / (time.perf_counter() - t0)
    h.close()
    ref_lp, ref_g = cb["row_logp"], cb["row_grad_sum"]
    e_lp = float(np.max(np.abs(lp - ref_lp) / np.maximum(np.abs(ref_lp), 1e-300)))
    scale = np.maximum(np.abs(ref_g), 1e-3 * np.max(np.abs(ref_g)))
    e_g = float(np.max(np.abs(g - ref_g) / scale))
    typ, k, _ = row_strata(dat[rows])
    return {"rows_checked": int(rows.shape[0]), "max_lattice_bits": int(k.max()),
            "logp_max_rel_err": e_lp, "grad_sum_max_rel_err": e_g,
            "parity_max_rel_err": max(e_lp, e_g), "tolerance": 1e-10, "ok": bool(max(e_lp, e_g) <= 1e-10),
            "against": "oracle/reference_restated.py on rows of THIS bench dataset (the stratified cpu_baseline sample)",
            "gpu_same_rows_patients_per_s": gpu_rate}


# ---- this repo's arm ---------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from metmhn_b200 import measure_fp64_tflops
    from metmhn_b200.sharded import ShardedEvaluator, class_weights

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL may print its version banner on stdout when the communicator is created; the contract is ONE JSON
        # line on stdout, so stdout points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    d = workload(args)
    dat, ep = d["dat"], d["eval_point"]
    t0 = time.perf_counter()
    ev = ShardedEvaluator(dat, rank=rank, world=world, device=local_rank, chunk_bytes=args.chunk_bytes,
                          rebalance=args.rebalance)
    setup_s = time.perf_counter() - t0
    w0, w1 = class_weights(ev.n_dat, ev.n_em, PERC_MET)
    par, out = ev.device_buffers()
    par.copy_(torch.from_numpy(ep))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        ev.step_device(w0, w1)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev.step_device(w0, w1)
        dev_ms += ev.handle.stats()["last_ms"]
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    dev_ms = max_over_ranks(dev_ms)
    clocks = sampler.stop() if sampler else {}
    result_dev = out.cpu().numpy().copy()
    launches = ev.handle.stats()["n_launches"] + (1 if world > 1 else 0)

    # end to end through the public API: host parameter vector in, host score + gradient out, every step
    for _ in range(2):
        ev.value_grad(ep, PERC_MET)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s_host, g_host = ev.value_grad(ep, PERC_MET)
    barrier()
    wall_e2e = max_over_ranks(time.perf_counter() - t0)
    # the two call paths (device-resident step, host API) must give the same numbers: reported, not asserted
    e2e_vs_device = float(max(abs(s_host - result_dev[0]) / abs(s_host),
                              np.max(np.abs(g_host - result_dev[1:])) / np.max(np.abs(g_host))))

    # per-class device times (CUDA events on the launching stream) for the roofline
    ev.handle.set_profile(True)
    ev.step_device(w0, w1)
    ev.step_device(w0, w1)
    st = ev.handle.stats()
    ev.handle.set_profile(False)
    cls = dict(st["class_ms"])
    detail = dict(cls)
    cls["contraction"] = cls.pop("stats") + cls.pop("finish") + cls.pop("pfin")     # marginal statistics + gradient contraction
    peaks, peak_src = load_peaks()
    states = st["states_value_grad"]
    # algorithmic bytes per class and step: forward writes y (8 B/state), adjoint writes x (8),
    # the contraction reads x and y (16); table setup is pure overhead (0).
    alg = {"setup": 0.0, "solve_fwd": 8.0 * states, "solve_adj": 8.0 * states, "contraction": 16.0 * states}
    kernels = {"setup": "k_setup, k_setup_wide, k_diag_prod",
               "solve_fwd": "k_solve_tile<fwd> (+ k_solve_big4 / k_solve_small* for the generic and small tiers)",
               "solve_adj": "k_solve_tile_adjb / k_solve_tile<adj> (+ k_solve_big4 / k_solve_small*)",
               "contraction": "k_stats_a/b, k_pf, k_finish"}
    # DRAM traffic of the tile solve kernels from the committed ncu capture (profiles/r2_final_solve_tile_ncu_full.txt,
    # 24 consecutive level launches of one chunk of 2^22..2^23-state pairs): forward 13.0 bytes per state (read +
    # write) against 8 algorithmic; adjoint with fused group-B statistics 17.6 (it also reads y)
    traffic_per_state = {"solve_fwd": 13.0, "solve_adj": 17.6}
    # the dominant KERNEL is the tile solve (forward and adjoint instantiations); the contraction class is several
    # kernels, each smaller
    dominant = "solve_adj" if cls["solve_adj"] >= cls["solve_fwd"] else "solve_fwd"
    if rank == 0:
        fp64_peak = measure_fp64_tflops(local_rank)
        top = dominant
        ach = alg[top] / (cls[top] * 1e-3) / 1e9 if cls[top] > 0 else 0.0
        ms_step = 1e3 * wall / args.steps
        # whole-step rooflines use this rank's share of the work and its device time
        t_rank = dev_ms / args.steps * 1e-3
        hbm_frac = st["alg_bytes"] / t_rank / 1e9 / peaks["hbm_gbs"]
        fp_frac = st["alg_flops"] / t_rank / 1e12 / fp64_peak
        line = {
            "metric": METRIC, "value": args.patients * args.steps / wall, "unit": "patients/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(base_config(args),
                       parallelism=f"patients sharded over {world} GPU(s) by LPT cost model + 1 all-reduce",
                       partition_rebalance={"rounds_allowed": args.rebalance if world > 1 else 0,
                                            "shard_ms_per_round": ev.rebalance_log,
                                            "what": "at construction (outside the timed region) every rank times its shard and "
                                                    "the rows are dealt again with capacities proportional to the measured speed"},
                       dataset_resident=True, dataset_upload_and_plan_s=round(setup_s, 3),
                       l2_policy="working set per step (x, y vectors of every patient, "
                                 f"{32.0 * states / 2 ** 30:.1f} GiB algorithmic) exceeds the 126 MB L2; no flush needed",
                       evals_per_s=args.steps / wall),
            "device_ms_per_step": dev_ms / args.steps,
            "clocks": clocks,
            "e2e": {"value": args.patients * args.steps / wall_e2e, "unit": "patients/s",
                    "h2d_bytes_per_step": int(ev.npar * 8), "d2h_bytes_per_step": int((ev.npar + 1) * 8),
                    "api": "ShardedEvaluator.value_grad(params_host, perc_met) -> (score, grad) host"},
            "gpu_launches": int(launches * args.steps),
            "e2e_vs_device_max_rel_diff": e2e_vs_device,
            "allreduce": ("in-library ncclAllReduce on the handle's stream (mmh_comm_init)" if ev.in_library_reduce
                          else "none (one GPU)"),
            "roofline": {"kernel": kernels[top], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / peaks["hbm_gbs"],
                         "traffic": (traffic_per_state[top] * states / 1e9) if top in traffic_per_state else None,
                         "traffic_unit": "GB per step, extrapolated from the bytes/state of the ncu capture in profiles/",
                         "algorithmic_GB_per_step": alg[top] / 1e9, "class_ms_serial": cls[top],
                         "peak_source": peak_src,
                         "note": "class time = CUDA events on the launching stream in the library's profile mode "
                                 "(chunks serialised on one stream); the timed steps overlap chunks on side streams",
                         "binding_resource": {
                             "what": "random 128-byte line reads through L2 (every state reads ~K/2 finished neighbours)",
                             "kernel_L2_sector_traffic_TBs": [5.8, 7.7], "kernel_L2_hit_rate": 0.6,
                             "ceiling_TBs": {"lines_in_L2_24_warps_per_SM": 6.7, "lines_in_HBM_24_warps_per_SM": 4.0},
                             "source": "profiles/r2_final_solve_tile_ncu_full.txt, profiles/r1_microbench_line_reads.txt "
                                       "(scripts/microbench/line_reads.cu); measured once per round, not in this run"}},
            "roofline_step": {"hbm": {"achieved_GBs": st["alg_bytes"] / t_rank / 1e9, "peak_GBs": peaks["hbm_gbs"], "frac": hbm_frac},
                              "fp64": {"achieved_TFLOPs": st["alg_flops"] / t_rank / 1e12, "peak_TFLOPs": fp64_peak,
                                       "frac": fp_frac, "peak_source": "independent-DFMA micro-kernel, this run"},
                              "binding": "hbm" if hbm_frac >= fp_frac else "fp64",
                              "states": states, "class_ms_serial": cls, "class_ms_detail": detail},
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_baseline(d, args.cpu_seconds)
            par = parity_vs_oracle(cb, dat, ep, local_rank)
            for key in ("row_index", "row_logp", "row_grad_sum", "seconds"):
                cb.pop(key, None)
            line["cpu_baseline"] = cb
            if par is not None:
                line["parity"] = par
                line["parity_max_rel_err"] = par["parity_max_rel_err"]
                line["rows_checked"] = par["rows_checked"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=25)
    ap.add_argument("--patients", type=int, default=100000)
    ap.add_argument("--chunk-bytes", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--ref-step-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--rebalance", type=int, default=3, help="measured-cost rebalancing rounds of the partition (N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    try:
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            run_ours(args, rank, local_rank, world)
    except BaseException:
        # torchrun swallows the traceback of a failing rank: say which rank, why, and what the library last reported
        import traceback
        msg = traceback.format_exc()
        try:
            from metmhn_b200 import _lib
            msg += f"mmh_last_error: {_lib.lib().mmh_last_error().decode('utf-8', 'replace')}\n"
        except Exception:
            pass
        sys.stderr.write(f"[bench rank {rank}/{world}] FAILED\n{msg}")
        sys.stderr.flush()
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"bench_error_rank{rank}.txt"), "w") as f:
                f.write(msg)
        except OSError:
            pass
        raise


if __name__ == "__main__":
    main()
