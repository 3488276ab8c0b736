"""Worker of tests/test_gpu_nccl.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Runs many `step_device` iterations on deliberately UNBALANCED shards (rank 0 gets the heavy rows, so the other ranks
run far ahead of it) and compares every iteration's reduced result with the single-GPU evaluation: the in-library
all-reduce is ordered with the evaluations on the handle's stream, so no iteration may see a stale or double sum."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from metmhn_b200 import Handle
    from metmhn_b200 import sharded
    from metmhn_b200.simulate import syn_v1

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = syn_v1(12, 3000, 12007, max_joint_bits=18)
    dat, ep = d["dat"], d["eval_point"]
    # unbalanced on purpose: the costliest tenth of the rows all go to rank 0
    cost = sharded.patient_cost(dat)
    heavy = np.argsort(-cost)[: dat.shape[0] // 10]
    assign = np.arange(dat.shape[0]) % world
    assign[heavy] = 0
    orig = sharded.partition
    sharded.partition = lambda dat_, world_: assign.astype(np.int32)
    try:
        ev = sharded.ShardedEvaluator(dat, rank=rank, world=world, device=local)
    finally:
        sharded.partition = orig
    assert ev.in_library_reduce
    w0, w1 = sharded.class_weights(ev.n_dat, ev.n_em, 0.65)
    ref = Handle(dat, device=local)
    par, out = ev.device_buffers()
    rng = np.random.default_rng(3)
    worst = 0.0
    for it in range(iters):
        p = ep + rng.normal(0, 0.02, ep.shape[0])            # same stream of parameters on every rank
        par.copy_(torch.from_numpy(p))
        torch.cuda.synchronize()
        ev.step_device(w0, w1)
        got = out.cpu().numpy()
        s, g = ref.value_grad(p, 0.65)
        want = np.concatenate([[s], g])
        worst = max(worst, float(np.max(np.abs(got - want)) / np.max(np.abs(want))))
    # host API path as well
    s2, g2 = ev.value_grad(ep, 0.65)
    s, g = ref.value_grad(ep, 0.65)
    worst = max(worst, abs(s2 - s) / abs(s), float(np.max(np.abs(g2 - g)) / np.max(np.abs(g))))
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "iters": iters, "max_rel_err": float(t.item())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
