"""Timed evaluations + per-class serial timings (library profile mode) for one synthetic workload.
usage: prof_classes.py n_events n_patients [seed_index]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metmhn_b200 import Handle
from metmhn_b200.simulate import syn_v1
n, nd = int(sys.argv[1]), int(sys.argv[2])
d = syn_v1(n, nd, 1000 * n + 3)
h = Handle(d['dat'])
ep = d['eval_point']
for _ in range(3):
    s, g = h.value_grad(ep, 0.65)
ts = []
for _ in range(5):
    t = time.perf_counter(); s, g = h.value_grad(ep, 0.65); ts.append(time.perf_counter() - t)
ms = 1e3 * float(np.median(ts))
h.set_profile(True)
h.value_grad(ep, 0.65)
s2, g2 = h.value_grad(ep, 0.65)
cls = {k: round(v, 2) for k, v in h.stats()['class_ms'].items()}
h.set_profile(False)
print(f"n={n} patients={nd} score={s:.15g} |g|={np.linalg.norm(g):.15g} same={s == s2 and np.array_equal(g, g2)}")
print(f"  {nd / ms * 1e3:.0f} patients/s  {ms:.2f} ms/eval  launches={h.stats()['n_launches']}  serial={sum(cls.values()):.1f} ms  {cls}")
