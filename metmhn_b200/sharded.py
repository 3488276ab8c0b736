"""Multi-GPU evaluation: one process per GPU, patients sharded by a cost model, one small all-reduce.

Patients are independent given the parameters (`regularized_optimization.py:187-254` only sums
per-patient results), so the dataset is partitioned once (longest-processing-time assignment on the
restricted state-space sizes) and every rank evaluates its shard with the GLOBAL class weights
(`:256-262` depend only on dataset counts).  The shard results then simply add up: the only exchange
is one all-reduce of 1 + (n+1)(n+3) doubles (5.8 KB at n = 25) over NCCL / NVLink.
"""
from __future__ import annotations

import heapq

import numpy as np


# Measured cost per lattice state in picoseconds on a B200, by patient kind and lattice size (scripts/calibrate_cost.py on the
# SYN-v1 n = 25 bench dataset, round 2): the big pair lattices run at the best rate, product-form single-tumour lattices
# cost twice as much per state (diagonal, weighted marginals), small lattices are latency-bound.  (bits >=, ps per state)
COST_PAIR = ((20, 37.0), (13, 66.0), (0, 485.0))
COST_PAIR_GENERIC = ((20, 195.0), (13, 590.0), (0, 485.0))     # fewer than 4 PT events or a tumour with more than 16: generic solve kernel
COST_SINGLE = ((17, 85.0), (13, 112.0), (9, 232.0), (0, 800.0))


def _rate(table, k):
    out = np.full(k.shape, table[-1][1])
    for bits, ps in reversed(table):
        out[k >= bits] = ps
    return out


def patient_cost(dat: np.ndarray) -> np.ndarray:
    """Work estimate per row: lattice states times the measured cost per state of the row's kind and size tier
    (the same table is in `row_cost`, metmhn_b200/csrc/metmhn_b200.cu, for the single-process multi-GPU path)."""
    dat = np.asarray(dat)
    n = (dat.shape[1] - 3) // 2
    typ = dat[:, -1]
    pt = dat[:, 0:2 * n:2].astype(np.int64).sum(axis=1)
    mt = dat[:, 1:2 * n:2].astype(np.int64).sum(axis=1)
    seed = dat[:, 2 * n].astype(np.int64)
    cost = np.ones(dat.shape[0])
    k1 = pt + seed
    s = (typ == 0) | (typ == 1)
    cost[s] = np.exp2(k1[s]) * _rate(COST_SINGLE, k1[s])
    s = typ == 2
    cost[s] = np.exp2(mt[s] + 1) * _rate(COST_SINGLE, mt[s] + 1)
    s = typ == 3
    kj = pt + mt
    generic = (pt[s] < 4) | (pt[s] > 16) | (mt[s] > 16)
    cost[s] = np.exp2(kj[s]) * np.where(generic, _rate(COST_PAIR_GENERIC, kj[s]), _rate(COST_PAIR, kj[s]))
    return cost


def partition(dat: np.ndarray, world: int, capacity=None) -> np.ndarray:
    """Rank of every row: greedy longest-processing-time assignment (deterministic).  `capacity[r]` (default: all equal)
    is the share of the work rank r should get; the next row goes to the rank with the smallest load / capacity."""
    cost = patient_cost(dat)
    if world <= 1:
        return np.zeros(dat.shape[0], dtype=np.int32)
    cap = np.ones(world) if capacity is None else np.asarray(capacity, dtype=np.float64)
    order = np.argsort(-cost, kind="stable")
    heap = [(0.0, r, 0.0) for r in range(world)]
    heapq.heapify(heap)
    out = np.empty(dat.shape[0], dtype=np.int32)
    for i in order:
        _, r, load = heapq.heappop(heap)
        out[i] = r
        load += cost[i]
        heapq.heappush(heap, (load / cap[r], r, load))
    return out


def rebalance_moves(assign: np.ndarray, cost: np.ndarray, times, damping: float = 0.7) -> np.ndarray:
    """Move rows from the ranks that measured slow to the ranks that measured fast (deterministic, incremental: everything
    else stays where it is, so the next measurement answers to the moves and not to a reshuffle).  A rank with time t and
    model load L should give away  damping * L * (1 - mean(t) / t)  of model cost; rows are taken largest first among those
    that fit the remaining surplus and go to the rank with the largest remaining deficit."""
    times = np.asarray(times, dtype=np.float64)
    world = times.shape[0]
    assign = assign.copy()
    load = np.array([cost[assign == r].sum() for r in range(world)])
    delta = damping * load * (times.mean() / times - 1.0)            # > 0: should receive
    delta -= delta.mean()
    deficit = np.maximum(delta, 0.0)
    for r in np.argsort(delta, kind="stable"):                       # donors, most overloaded first
        surplus = -delta[r]
        if surplus <= 0:
            break
        rows = np.nonzero(assign == r)[0]
        rows = rows[np.argsort(-cost[rows], kind="stable")]
        for i in rows:
            if surplus <= 0 or deficit.max() <= 0:
                break
            c = cost[i]
            if c > surplus:
                continue
            q = int(np.argmax(deficit))
            if c > 1.25 * deficit[q]:
                continue
            assign[i] = q
            deficit[q] -= c
            surplus -= c
    return assign


def class_weights(n_dat: int, n_em: float, perc_met: float):
    """(w_type0, w_other) = (1, w) / n_full of regularized_optimization.py:256-262."""
    n_nm = n_dat - n_em
    w = perc_met * n_nm / ((1.0 - perc_met) * n_em) if n_em * n_nm != 0 else 1.0
    n_full = w * n_em + n_nm
    return 1.0 / n_full, w / n_full


class ShardedEvaluator:
    """value / value_grad of a dataset sharded over the ranks of a torch.distributed group.

    Every rank constructs it with the full `dat` (or, with `rows=`, only its own rows plus the global
    counts).  `local_eval(params, w0, w1, want_grad) -> np.ndarray[1 + npar]` is the shard evaluator; by
    default it is the CUDA handle of this rank's shard."""

    def __init__(self, dat, rank=0, world=1, device=0, group=None, local_eval=None, chunk_bytes=0, rebalance=0):
        dat = np.ascontiguousarray(np.asarray(dat), dtype=np.int8)
        self.rank, self.world, self.group = rank, world, group
        self.n_mut = (dat.shape[1] - 3) // 2
        self.n_tot = self.n_mut + 1
        self.npar = self.n_tot * (self.n_tot + 2)
        self.n_dat = dat.shape[0]
        self.n_em = float(dat[:, 2 * self.n_mut].astype(np.int64).sum())
        self.assign = partition(dat, world)
        self.shard = np.ascontiguousarray(dat[self.assign == rank])
        self.shard_cost = float(patient_cost(self.shard).sum()) if self.shard.shape[0] else 0.0
        self.handle = None
        self._dev_out = None
        self._dev_par = None
        self.in_library_reduce = False
        self.rebalance_log = []
        if local_eval is None:
            from ._lib import Handle
            self.handle = Handle(self.shard, device=device, chunk_bytes=chunk_bytes)
            for _ in range(rebalance if world > 1 else 0):
                if not self._rebalance_once(dat, device, chunk_bytes):
                    break
            local_eval = self._cuda_eval
            if world > 1:
                self._attach_comm()
        self.local_eval = local_eval

    def _rebalance_once(self, dat, device, chunk_bytes):
        """Measured-cost rebalancing (optional, at construction): every rank times its shard, the times are gathered, and
        the rows are dealt again with capacities proportional to the measured speed (model cost per millisecond) of every
        rank.  The static cost model equalises lattice states by kind and size; what it cannot see is how a shard's
        particular lattice shapes pack into level launches (measured: +-7 % between shards of equal model cost)."""
        import torch.distributed as dist
        from ._lib import Handle
        x = np.zeros(self.npar)
        for _ in range(2):
            self.handle.eval_weighted(x, 1.0, 1.0)
        ms = []
        for _ in range(3):
            self.handle.eval_weighted(x, 1.0, 1.0)
            ms.append(float(self.handle.stats()["last_ms"]))
        times = [None] * self.world
        dist.all_gather_object(times, min(ms), group=self.group)
        times = np.asarray(times, dtype=np.float64)
        self.rebalance_log.append([round(float(t), 3) for t in times])
        if times.max() <= 1.015 * times.mean():
            return False
        cost = patient_cost(dat)
        self.assign = rebalance_moves(self.assign, cost, times)
        self.shard = np.ascontiguousarray(dat[self.assign == self.rank])
        self.shard_cost = float(cost[self.assign == self.rank].sum()) if self.shard.shape[0] else 0.0
        self.handle.close()
        self.handle = Handle(self.shard, device=device, chunk_bytes=chunk_bytes)
        return True

    def _attach_comm(self):
        """Give the handle its own NCCL communicator (id from rank 0, shipped through the torch.distributed group
        that launched the ranks): the all-reduce then runs inside the library on the handle's stream, ordered with
        the evaluation that produced the result and with the next one that overwrites it."""
        import torch.distributed as dist
        from ._lib import nccl_unique_id
        box = [nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=self.group)
        self.handle.comm_init(box[0], self.world, self.rank)
        self.in_library_reduce = True

    def _cuda_eval(self, params, w0, w1, want_grad):
        s, g = self.handle.eval_weighted(params, w0, w1, want_grad=want_grad)
        return np.concatenate([[s], g]) if want_grad else np.array([s])

    def _reduce(self, vec):
        if self.world <= 1 or self.in_library_reduce:
            return vec
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(vec))
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def value_grad(self, params, perc_met):
        w0, w1 = class_weights(self.n_dat, self.n_em, perc_met)
        out = self._reduce(self.local_eval(np.asarray(params, dtype=np.float64), w0, w1, True))
        return float(out[0]), out[1:]

    def value(self, params, perc_met):
        w0, w1 = class_weights(self.n_dat, self.n_em, perc_met)
        out = self._reduce(self.local_eval(np.asarray(params, dtype=np.float64), w0, w1, False))
        return float(out[0])

    # ---- device-resident step used by bench.py: no host copies inside ---------------------------------
    def device_buffers(self):
        import torch
        if self._dev_out is None:
            dev = torch.device("cuda", self.handle.device)
            self._dev_out = torch.zeros(self.npar + 1, dtype=torch.float64, device=dev)
            self._dev_par = torch.zeros(self.npar, dtype=torch.float64, device=dev)
        return self._dev_par, self._dev_out

    def step_device(self, w0, w1, want_grad=True):
        """One evaluation with parameters and result resident in HBM.  With world > 1 the library ends the
        evaluation with its own ncclAllReduce on the handle's stream, so `out` is never touched by two streams."""
        par, out = self.device_buffers()
        self.handle.eval_device(par.data_ptr(), w0, w1, out.data_ptr(), want_grad)
        self.handle.sync()
        return out
