"""CPU tests of the multi-GPU host logic: cost-model partition and the world-size-2 reduction (gloo).
The shard evaluator is injected (CPU oracle) so that the plumbing is exercised without a GPU."""
import os
import socket

import numpy as np
import pytest

from metmhn_b200.sharded import ShardedEvaluator, class_weights, partition, patient_cost, rebalance_moves
from metmhn_b200.simulate import syn_v1


def test_partition_is_balanced_and_complete():
    d = syn_v1(12, 2000, 12)
    dat = d["dat"]
    for world in (2, 4, 8):
        a = partition(dat, world)
        assert a.shape[0] == dat.shape[0] and set(np.unique(a)) == set(range(world))
        cost = patient_cost(dat)
        loads = np.array([cost[a == r].sum() for r in range(world)])
        # LPT guarantee: max load <= mean + largest item
        assert loads.max() <= loads.mean() + cost.max() + 1e-9
    assert np.array_equal(partition(dat, 4), partition(dat, 4))      # deterministic
    # capacities (measured-cost rebalancing): the loads follow the requested shares
    a = partition(dat, 3, capacity=[0.5, 0.3, 0.2])
    cost = patient_cost(dat)
    loads = np.array([cost[a == r].sum() for r in range(3)])
    assert np.allclose(loads / loads.sum(), [0.5, 0.3, 0.2], atol=0.02)


def test_rebalance_moves_shift_load_from_slow_to_fast_ranks():
    d = syn_v1(12, 2000, 12)
    dat = d["dat"]
    cost = patient_cost(dat)
    a0 = partition(dat, 4)
    load0 = np.array([cost[a0 == r].sum() for r in range(4)])
    times = np.array([1.10, 1.00, 0.95, 0.95])                     # rank 0 measured 10 % slow
    a1 = rebalance_moves(a0, cost, times, damping=1.0)
    load1 = np.array([cost[a1 == r].sum() for r in range(4)])
    assert load1[0] < load0[0] and load1[2] > load0[2] and load1[3] > load0[3]
    want = load0 * times.mean() / times
    assert np.allclose(load1 / load1.sum(), want / want.sum(), atol=0.02)
    assert np.array_equal(a1, rebalance_moves(a0, cost, times, damping=1.0))           # deterministic
    assert np.array_equal(rebalance_moves(a0, cost, np.ones(4)), a0)                   # balanced: nothing moves


def test_library_cost_model_equals_the_python_one():
    """`row_cost` of the library (partition of mmh_multi_create) and `patient_cost` (ShardedEvaluator) are the same table."""
    import ctypes as C
    import os
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "metmhn_b200", "libmetmhn_b200.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    L = C.CDLL(lib)
    L.mmh_row_cost.restype = C.c_double
    L.mmh_row_cost.argtypes = [C.c_void_p, C.c_int]
    for n, nd, seed in ((12, 400, 5), (20, 300, 7), (25, 300, 9)):
        dat = np.ascontiguousarray(syn_v1(n, nd, seed)["dat"], dtype=np.int8)
        want = patient_cost(dat)
        got = np.array([L.mmh_row_cost(dat[r].ctypes.data, n) for r in range(dat.shape[0])])
        assert np.array_equal(got, want)


def test_class_weights_match_reference_formula():
    w0, w1 = class_weights(100, 80, 0.65)
    w = 0.65 * 20 / (0.35 * 80)
    assert abs(w1 / w0 - w) < 1e-15 and abs(1 / w0 - (w * 80 + 20)) < 1e-12
    assert class_weights(10, 0, 0.3) == (0.1, 0.1)                    # n_em * n_nm == 0 -> w = 1


def _oracle_local_eval(shard):
    from oracle import lattice_direct as ld

    def f(params, w0, w1, want_grad):
        n = (shard.shape[1] - 3) // 2
        n_tot = n + 1
        th, dp, dm = params[:n_tot ** 2].reshape(n_tot, n_tot), params[n_tot ** 2:n_tot ** 2 + n_tot], params[n_tot ** 2 + n_tot:]
        out = np.zeros(1 + n_tot * (n_tot + 2))
        for r in shard:
            o = ld.patient_value_grad(th, dp, dm, r)
            if o is None:
                continue
            w = w0 if o[0] else w1
            out[0] += w * o[1]
            out[1:] += w * np.concatenate([o[2].ravel(), o[3], o[4]])
        return out if want_grad else out[:1]
    return f


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = syn_v1(5, 40, 55)
    ev = ShardedEvaluator(d["dat"], rank=rank, world=world, local_eval=lambda *a: None)
    ev.local_eval = _oracle_local_eval(ev.shard)
    s, g = ev.value_grad(d["eval_point"], 0.65)
    v = ev.value(d["eval_point"], 0.65)
    q.put((rank, s, g, v, ev.shard.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import lattice_direct as ld
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    d = syn_v1(5, 40, 55)
    ep = d["eval_point"]
    s0, g0, a0, b0 = ld.score_and_grad(ep[:36].reshape(6, 6), ep[36:42], ep[42:], d["dat"], 0.65)
    ref = np.concatenate([g0.ravel(), a0, b0])
    assert res[0][4] + res[1][4] == 40
    for _, s, g, v, _ in res:
        assert abs(s - s0) <= 1e-12 * abs(s0) and abs(v - s0) <= 1e-12 * abs(s0)
        assert np.abs(g - ref).max() <= 1e-12 * np.abs(ref).max()


def _cv_worker(rank, world, port, q):
    import torch.distributed as dist
    from metmhn_b200.utility import cross_val_distributed
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = syn_v1(4, 60, 44)
    fake = lambda train, test, lam: float(train.shape[0] * 1000 + test[:, :-2].sum() + lam)     # deterministic stand-in for a fit
    runs = cross_val_distributed(d["dat"], None, np.array([0.1, 0.2, 0.3]), 4, 0.65, rank=rank, world=world, fit_and_score=fake)
    q.put((rank, runs))
    dist.barrier()
    dist.destroy_process_group()


def test_cross_val_jobs_are_distributed_and_gathered():
    """BASELINE config 5 (fold x lambda sweep over the GPUs of a box): job distribution + result gathering on gloo."""
    import torch.multiprocessing as mp
    from metmhn_b200.utility import cross_val_distributed
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cv_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    d = syn_v1(4, 60, 44)
    fake = lambda train, test, lam: float(train.shape[0] * 1000 + test[:, :-2].sum() + lam)
    single = cross_val_distributed(d["dat"], None, np.array([0.1, 0.2, 0.3]), 4, 0.65, fit_and_score=fake)
    assert single.shape == (4, 3) and (single != 0).all()
    for _, runs in res:
        assert np.array_equal(runs, single)
