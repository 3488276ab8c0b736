"""Step time of ONE rank's shard (rank 0 of `world`) on one GPU for several chunk budgets x side streams: what limits the
strong-scaling curve is the per-rank chain of thin launches, which this emulates without the other GPUs.
usage: sweep_shard.py n patients world chunk_bytes,... streams,..."""
import sys, os, subprocess
code = r'''
import sys, os, time
sys.path.insert(0, %r)
import numpy as np
from metmhn_b200 import Handle
from metmhn_b200.simulate import syn_v1
from metmhn_b200.sharded import partition
n, nd, world, chunk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(float(sys.argv[4]))
d = syn_v1(n, nd, 1000 * n + 3)
dat = d['dat']
shard = np.ascontiguousarray(dat[partition(dat, world) == 0])
h = Handle(shard, chunk_bytes=chunk)
ep = d['eval_point']
for _ in range(3): h.eval_weighted(ep, 1.0, 1.0)
ms = []
for _ in range(6):
    h.eval_weighted(ep, 1.0, 1.0); ms.append(h.stats()['last_ms'])
print('RESULT', min(ms), h.stats()['n_chunks'], h.stats()['n_launches'])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n, nd, world = sys.argv[1], sys.argv[2], sys.argv[3]
for chunk in sys.argv[4].split(","):
    for ns in sys.argv[5].split(","):
        out = subprocess.run([sys.executable, "-c", code, n, nd, world, chunk], env=dict(os.environ, MMH_STREAMS=ns), capture_output=True, text=True)
        res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        print(f"world={world} chunk={float(chunk)/2**20:.0f}MiB streams={ns}: {res[0] if res else out.stderr[-300:]}", flush=True)
