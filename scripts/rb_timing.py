"""Phase timing of k_solve_rb (library built with -DRB_TIMING, MMH_LIB=.../lib_rbt.so): cycles per block and phase."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metmhn_b200 import Handle, _lib
from metmhn_b200.simulate import syn_v1
n, nd = int(sys.argv[1]), int(sys.argv[2])
d = syn_v1(n, nd, 1000 * n + 3)
h = Handle(d['dat'])
ep = d['eval_point']
L = C.CDLL(_lib.LIB_PATH)
buf = (C.c_ulonglong * 16)()
h.value_grad(ep, 0.65)
L.mmh_debug_rb_timing(buf, 1)
h.value_grad(ep, 0.65)
L.mmh_debug_rb_timing(buf, 1)
t = list(buf)
nblk = t[10] + t[11]
print("CTAs fwd/adj", t[8], t[9], "blocks fwd/adj", t[10], t[11])
names = ["tables+sync", "context", "phase1", "phase2", "phase3", "stats"]
for i, nm in enumerate(names):
    print(f"  {nm:12s} {t[i] / max(nblk, 1):10.0f} cycles per block")
print("  total        %10.0f cycles per block" % (sum(t[:6]) / max(nblk, 1)))
