"""Derive the int8 data matrix of the reference's own runnable workload (examples/analysis.py on data/luad) and
store it as a fixture, so that the workload can run where /root/reference is absent (the GPU box).
Run in the build container:  python tests/golden/make_luad_fixture.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from metmhn_b200.utility import read_events_csv  # noqa: E402

ref = os.environ.get("METMHN_REFERENCE_ROOT", "/root/reference")
dat, names = read_events_csv(os.path.join(ref, "data/luad/G14_LUAD_Events.csv"),
                             os.path.join(ref, "data/luad/G14_LUAD_sampleSelection.csv"))
typ, cnt = np.unique(dat[:, -1], return_counts=True)
print(dat.shape, dict(zip(typ.tolist(), cnt.tolist())), dict(zip(*np.unique(dat[dat[:, -1] == 3, -2], return_counts=True))))
np.savez_compressed(os.path.join(HERE, "luad_dat.npz"), dat=dat, events=np.array(names))
