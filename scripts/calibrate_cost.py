"""Measured cost per lattice state by patient kind and size tier (the partition's cost model, metmhn_b200/sharded.py, is
calibrated with these): evaluates subsets of the bench dataset alone.   usage: calibrate_cost.py n patients"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metmhn_b200 import Handle
from metmhn_b200.simulate import syn_v1
n, nd = int(sys.argv[1]), int(sys.argv[2])
d = syn_v1(n, nd, 1000 * n + 3)
dat, ep = d['dat'], d['eval_point']
typ = dat[:, -1]
pt = dat[:, 0:2 * n:2].astype(np.int64).sum(axis=1); mt = dat[:, 1:2 * n:2].astype(np.int64).sum(axis=1); seed = dat[:, 2 * n].astype(np.int64)
k = np.where(typ == 3, pt + mt, np.where(typ == 2, mt + 1, pt + seed))
gen = (typ == 3) & ((pt < 4) | (pt > 16) | (mt > 16))          # pairs on the generic solve kernel
groups = {"generic pairs K>=20": gen & (k >= 20), "generic pairs 13<=K<20": gen & (k >= 13) & (k < 20),
          "tiled pairs K>=20": (typ == 3) & ~gen & (k >= 20), "tiled pairs 13<=K<20": (typ == 3) & ~gen & (k >= 13) & (k < 20),
          "pairs K>=20": (typ == 3) & (k >= 20), "pairs 13<=K<20": (typ == 3) & (k >= 13) & (k < 20), "pairs K<13": (typ == 3) & (k < 13),
          "type0/1 K>=17": (typ <= 1) & (k >= 17), "type0/1 13<=K<17": (typ <= 1) & (k >= 13) & (k < 17), "type0/1 9<=K<13": (typ <= 1) & (k >= 9) & (k < 13), "type0/1 K<9": (typ <= 1) & (k < 9),
          "type2 K>=17": (typ == 2) & (k >= 17), "type2 13<=K<17": (typ == 2) & (k >= 13) & (k < 17), "type2 9<=K<13": (typ == 2) & (k >= 9) & (k < 13), "type2 K<9": (typ == 2) & (k < 9)}
for name, m in groups.items():
    sub = np.ascontiguousarray(dat[m])
    if sub.shape[0] == 0: continue
    h = Handle(sub)
    for _ in range(3): h.eval_weighted(ep, 1.0, 1.0)
    ms = []
    for _ in range(5):
        h.eval_weighted(ep, 1.0, 1.0); ms.append(h.stats()['last_ms'])
    st = h.stats()['states_value_grad']
    print(f"{name:18s} rows {sub.shape[0]:6d}  states {st:.3e}  {min(ms):8.3f} ms  {1e9 * min(ms) / st:8.1f} ps/state  {1e3 * min(ms) / sub.shape[0]:8.3f} us/row", flush=True)
    h.close()
