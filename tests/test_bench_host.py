"""Host-side pieces of bench.py that need no GPU: the bounded CPU sample of the reference arm and the JSON contract
keys of the reference line."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_cpu_baseline_sample_is_bounded_and_counts_rows():
    import bench
    from metmhn_b200.simulate import syn_v1
    d = syn_v1(8, 200, 8003)
    out = bench.cpu_baseline(d, 0, rows_limit=24)
    assert out["rows"] == 24 and out["kind"] == "port" and out["cores"] >= 1
    assert out["value"] > 0 and out["states_per_s"] > 0
    idx, bits = bench.cpu_sample_rows(d["dat"], 11)
    assert np.all(bits[idx] <= 11) and np.array_equal(out["row_index"], idx[:24])


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "6", "--patients", "300",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "patients/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
