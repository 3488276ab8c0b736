"""Config C5 of BASELINE.json: 5-fold x 5-lambda cross-validation (n = 20 events, 20 000 patients,
lambda = 10^linspace(-3.5, -2.5, 5), data_analysis.ipynb cell 12) with the 25 independent fits dealt to the ranks of a
torch.distributed group, one rank per GPU (metmhn_b200.utility.cross_val_distributed).  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/c5_cv.py
Rank 0 prints one JSON line: wall time of the sweep, the (fold x lambda) score table, the chosen lambda and a parity check of
one (fold, lambda) cell against the serial path on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    import torch
    import torch.distributed as dist
    import metmhn_b200 as mm
    from metmhn_b200.simulate import syn_v1
    from metmhn_b200.utility import cross_val_distributed, indep
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n_dat = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    folds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    nlam = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    d = syn_v1(20, n_dat, 20005)
    dat = d["dat"]
    lams = 10.0 ** np.linspace(-3.5, -2.5, nlam)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    runs = cross_val_distributed(dat, mm.symmetric_penal, lams, folds, 0.65, seed=42, rank=rank, world=world, device=local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    if rank == 0:
        # parity of one cell with the serial path (same shuffle, same fold, same lambda, this GPU)
        shuffled = dat[np.random.Generator(np.random.PCG64(42)).permutation(dat.shape[0])]
        batch = int(np.ceil(dat.shape[0] / folds))
        fold, i = folds - 1, nlam // 2
        start, stop = batch * fold, min(batch * (fold + 1), dat.shape[0])
        train = np.ascontiguousarray(np.concatenate([shuffled[:start], shuffled[stop:]]))
        test = np.ascontiguousarray(shuffled[start:stop])
        t1 = time.perf_counter()
        th0, dp0, dm0 = indep(train)
        th, dp, dm = mm.learn_mhn(th0, dp0, dm0, train, 0.65, mm.symmetric_penal, float(lams[i]), opt_v=False)
        one = float(mm.score(th, dp, dm, test, 0.65))
        t_one = time.perf_counter() - t1
        mean = runs.mean(axis=0)
        print(json.dumps({"config": f"C5: SYN-v1 n=20, {n_dat} patients, {folds} folds x {nlam} lambdas = {folds * nlam} fits over {world} GPU(s)",
                          "wall_s": wall, "fits": folds * nlam, "n_gpus": world, "lambdas": lams.tolist(), "scores": runs.tolist(),
                          "mean_score_per_lambda": mean.tolist(), "chosen_lambda": float(lams[int(np.argmax(mean))]),
                          "serial_cell": {"fold": fold, "lambda_index": i, "score": one, "distributed_score": float(runs[fold, i]),
                                          "abs_diff": abs(one - float(runs[fold, i])), "seconds_for_one_fit_and_score": t_one}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
