"""Multi-GPU path on real GPUs over NCCL (skipped on a one-GPU box): the sharded `step_device` with the in-library
all-reduce equals the single-GPU `value_grad` over many iterations with rank skew; the one-process multi-device
entry points (`mmh_multi_*`) equal it as well."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_step_device_equals_single_gpu_over_many_steps():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "nccl_worker.py"), "40"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-4000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["iters"] == 40
    assert line["max_rel_err"] <= 1e-12, line


def test_one_process_multi_device_handle():
    from metmhn_b200 import Handle
    from metmhn_b200._lib import MultiHandle
    from metmhn_b200.simulate import syn_v1
    n = _ngpu()
    d = syn_v1(10, 1500, 10011, max_joint_bits=16)
    ref = Handle(d["dat"])
    s, g = ref.value_grad(d["eval_point"], 0.65)
    for devs in ([0], list(range(min(n, 4)))):
        m = MultiHandle(d["dat"], devs)
        for _ in range(3):
            s2, g2 = m.value_grad(d["eval_point"], 0.65)
            assert abs(s2 - s) <= 1e-12 * abs(s)
            assert np.max(np.abs(g2 - g)) <= 1e-12 * np.max(np.abs(g))
        assert abs(m.value(d["eval_point"], 0.65) - s) <= 1e-12 * abs(s)
        m.close()
