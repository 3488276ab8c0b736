"""Phase timing of k_solve_rb for ONE pair (unloaded latencies): usage rb_timing_one.py KA KB"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metmhn_b200 import Handle, _lib
ka, kb = int(sys.argv[1]), int(sys.argv[2])
n = 26
rng = np.random.default_rng(1)
th = rng.normal(0.0, 0.3, (n + 1, n + 1)); th[np.arange(n + 1), np.arange(n + 1)] = rng.normal(-1.0, 0.5, n + 1)
params = np.concatenate([th.ravel(), rng.normal(0, 0.3, n + 1), rng.normal(0, 0.3, n + 1)])
row = np.zeros((1, 2 * n + 3), dtype=np.int8)
row[0, 0:2 * ka:2] = 1                       # PT events 0..ka-1
row[0, 2 * ka + 1:2 * (ka + kb) + 1:2] = 1   # MT events ka..ka+kb-1
row[0, 2 * n] = 1; row[0, -2:] = (1, 3)
h = Handle(row)
L = C.CDLL(_lib.LIB_PATH)
buf = (C.c_ulonglong * 16)()
h.value_grad(params, 0.65); L.mmh_debug_rb_timing(buf, 1)
h.value_grad(params, 0.65); L.mmh_debug_rb_timing(buf, 1)
t = list(buf); nblk = t[10] + t[11]
print("KA", ka, "KB", kb, "CTAs fwd/adj", t[8], t[9], "blocks", t[10], t[11])
for i, nm in enumerate(["tables+sync", "context", "phase1", "phase2", "phase3", "stats"]):
    print(f"  {nm:12s} {t[i] / max(nblk, 1):10.0f} cycles per block")
print("  per block: thread 0 prepare %.0f solve %.0f | thread 255 prepare %.0f solve %.0f" % tuple(t[i] / max(nblk, 1) for i in (12, 13, 14, 15)))
