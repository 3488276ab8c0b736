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
    from oracle import reference_restated as rr
    th, dp, dm, row = args
    t = time.perf_counter()
    rr.patient_value_grad(th, dp, dm, row, want_grad=True)
    return time.perf_counter() - t


def cpu_sample_rows(dat, max_bits):
    n = (dat.shape[1] - 3) // 2
    typ = dat[:, -1]
    bits = np.where(typ == 3, dat[:, :2 * n + 1].astype(np.int64).sum(axis=1),
                    np.where(typ == 2, dat[:, 1:2 * n:2].astype(np.int64).sum(axis=1) + 1,
                             dat[:, 0:2 * n + 1:2].astype(np.int64).sum(axis=1)))
    return np.nonzero(bits <= max_bits)[0], bits


def cpu_baseline(d, budget_s, max_bits=11, rows_limit=None):
    """Time the reference algorithm (oracle/reference_restated.py: per-event Kronecker shuffles, (k+1) Jacobi
    sweeps) on all host cores over a bounded sample: rows in dataset order whose lattice has <= 2^max_bits
    states, as many as fit in the time budget."""
    import multiprocessing as mp
    dat = d["dat"]
    n = (dat.shape[1] - 3) // 2
    ep = d["eval_point"]
    n_tot = n + 1
    th, dp, dm = ep[:n_tot * n_tot].reshape(n_tot, n_tot), ep[n_tot * n_tot:n_tot * (n_tot + 1)], ep[n_tot * (n_tot + 1):]
    idx, bits = cpu_sample_rows(dat, max_bits)
    share_rows = idx.shape[0] / dat.shape[0]
    if rows_limit:
        idx = idx[:rows_limit]
    cores = os.cpu_count() or 1
    done = 0
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        it = pool.imap(_cpu_one, ((th, dp, dm, dat[i]) for i in idx), chunksize=4)
        for _ in it:
            done += 1
            if rows_limit is None and time.perf_counter() - t0 > budget_s:
                break
        pool.terminate()
    el = time.perf_counter() - t0
    used = idx[:done]
    states = float(np.exp2(bits[used]).sum())
    return {"value": done / el, "unit": "patients/s", "cores": cores, "kind": "port",
            "sample": f"first {done} rows (dataset order) with lattice <= 2^{max_bits} states "
                      f"({100 * share_rows:.1f}% of rows qualify); reference algorithm = oracle/reference_restated.py "
                      f"(NumPy restatement: JAX is not installable here), multiprocessing over rows, {el:.1f} s",
            "rows": int(done), "seconds": el, "states_per_s": states / el, "row_index": used}


def run_reference(args, rank, world):
    if rank != 0:
        return
    d = workload(args)
    # calibrate the sample so that one step is ~8 s of all-core CPU work
    cal = cpu_baseline(d, 6.0)
    rows = max(8, int(cal["value"] * 8.0))
    for _ in range(args.warmup):
        cpu_baseline(d, 0, rows_limit=max(8, rows // 8))
    t = []
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(d, 0, rows_limit=rows)
        t.append(last["seconds"])
    val = last["rows"] * len(t) / sum(t)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "patients/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / len(t),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(base_config(args), cpu_step="bounded sample of the workload's rows, see cpu_baseline.sample"),
            "cpu_baseline": {"value": val, "unit": "patients/s", "cores": last["cores"], "kind": "port",
                             "sample": last["sample"]},
            "e2e": {"value": val, "unit": "patients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    ev = ShardedEvaluator(dat, rank=rank, world=world, device=local_rank, chunk_bytes=args.chunk_bytes)
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
    assert abs(s_host - result_dev[0]) <= 1e-12 * abs(s_host)

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
               "contraction": "k_stats_a/b, k_pf_lo/hi, k_finish"}
    # DRAM traffic of the tile solve kernels from the committed ncu capture (profiles/r1_final_solve_tile_ncu_full.txt,
    # 24 consecutive level launches of one chunk of 2^22..2^23-state pairs): forward 12.9 bytes per state (read +
    # write) against 8 algorithmic; adjoint with fused group-B statistics 16.9 (it also reads y)
    traffic_per_state = {"solve_fwd": 12.9, "solve_adj": 16.9}
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
                             "kernel_L2_sector_traffic_TBs": [6.0, 6.4], "kernel_L2_hit_rate": 0.6,
                             "ceiling_TBs": {"lines_in_L2_24_warps_per_SM": 6.7, "lines_in_HBM_24_warps_per_SM": 4.0},
                             "source": "profiles/r1_final_solve_tile_ncu_full.txt, profiles/r1_microbench_line_reads.txt "
                                       "(scripts/microbench/line_reads.cu); measured once per round, not in this run"}},
            "roofline_step": {"hbm": {"achieved_GBs": st["alg_bytes"] / t_rank / 1e9, "peak_GBs": peaks["hbm_gbs"], "frac": hbm_frac},
                              "fp64": {"achieved_TFLOPs": st["alg_flops"] / t_rank / 1e12, "peak_TFLOPs": fp64_peak,
                                       "frac": fp_frac, "peak_source": "independent-DFMA micro-kernel, this run"},
                              "binding": "hbm" if hbm_frac >= fp_frac else "fp64",
                              "states": states, "class_ms_serial": cls, "class_ms_detail": detail},
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_baseline(d, args.cpu_seconds)
            rows = cb.pop("row_index")
            from metmhn_b200 import Handle
            hs = Handle(np.ascontiguousarray(dat[rows]), device=local_rank)
            hs.value_grad(ep, PERC_MET)
            t0 = time.perf_counter()
            for _ in range(5):
                hs.value_grad(ep, PERC_MET)
            cb["gpu_same_sample_value"] = rows.shape[0] * 5 / (time.perf_counter() - t0)
            hs.close()
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
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
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
