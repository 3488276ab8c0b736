import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metmhn_b200 import Handle
from metmhn_b200.simulate import syn_v1
n, nd = int(sys.argv[1]), int(sys.argv[2])
chunk = int(float(sys.argv[3])) if len(sys.argv) > 3 else 0
t = time.time(); d = syn_v1(n, nd, 1000 * n + 3); print('gen', time.time() - t, flush=True)
t = time.time(); h = Handle(d['dat'], chunk_bytes=chunk); print('create', time.time() - t, flush=True)
st = h.stats(); print({k: v for k, v in st.items() if k != 'k_hist'}, flush=True)
for i in range(int(os.environ.get("EVALS", "3"))):
    t = time.time(); s, g = h.value_grad(d['eval_point'], 0.65); dt = time.time() - t
    st = h.stats()
    print('eval', i, 'wall', dt, 'dev_ms', st['last_ms'], 'launches', st['n_launches'], 'score', s, 'patients/s', nd / dt, flush=True)
if os.environ.get("VALUE_ONLY", "1") == "1":
    t = time.time(); v = h.value(d['eval_point'], 0.65); print('value only', time.time() - t, v, h.stats()['last_ms'])
