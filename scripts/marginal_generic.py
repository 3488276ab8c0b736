import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from metmhn_b200 import Handle
from metmhn_b200.simulate import syn_v1
for n, nd in ((20, 10000), (25, 100000)):
    d = syn_v1(n, nd, 1000 * n + 3)
    dat, ep = d['dat'], d['eval_point']
    typ = dat[:, -1]
    pt = dat[:, 0:2 * n:2].astype(np.int64).sum(axis=1); mt = dat[:, 1:2 * n:2].astype(np.int64).sum(axis=1)
    gen = (typ == 3) & ((pt < 4) | (pt > 16) | (mt > 16)) & (pt + mt >= 13)
    wide = (typ == 3) & ((pt > 16) | (mt > 16))
    for name, m in (("all", np.ones(len(dat), bool)), ("without generic pairs", ~gen), ("without wide pairs", ~wide)):
        sub = np.ascontiguousarray(dat[m])
        h = Handle(sub)
        for _ in range(3): h.eval_weighted(ep, 1.0, 1.0)
        ms = []
        for _ in range(5):
            h.eval_weighted(ep, 1.0, 1.0); ms.append(h.stats()['last_ms'])
        print(n, name, 'rows', sub.shape[0], 'removed', int((~m).sum()), 'states %.3e' % h.stats()['states_value_grad'], '%.2f ms' % min(ms), flush=True)
        h.close()
