"""Evaluation time of every rank's shard of a `world`-way partition, each measured alone on one GPU, with the states and the
cost-model value of the shard: shows how well the partition's cost model matches the measured cost.
usage: shard_times.py n patients world"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metmhn_b200 import Handle
from metmhn_b200.simulate import syn_v1
from metmhn_b200.sharded import partition, patient_cost
n, nd, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = syn_v1(n, nd, 1000 * n + 3)
dat, ep = d['dat'], d['eval_point']
assign = partition(dat, world)
cost = patient_cost(dat)
for r in range(world):
    shard = np.ascontiguousarray(dat[assign == r])
    h = Handle(shard)
    for _ in range(3): h.eval_weighted(ep, 1.0, 1.0)
    ms = []
    for _ in range(5):
        h.eval_weighted(ep, 1.0, 1.0); ms.append(h.stats()['last_ms'])
    st = h.stats()
    typ = shard[:, -1]
    print(f"rank {r}: {min(ms):7.2f} ms  rows {shard.shape[0]:6d}  states {st['states_value_grad']:.3e}  model cost {cost[assign == r].sum():.3e}  paired rows {(typ == 3).sum()}", flush=True)
    h.close()
