"""Host-side pieces of bench.py that need no GPU: the bounded CPU sample of the reference arm and the JSON contract
keys of the reference line."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_cpu_baseline_is_stratified_extrapolated_and_keeps_the_oracle_results():
    import bench
    from metmhn_b200.simulate import syn_v1
    from oracle import reference_restated as rr
    d = syn_v1(8, 200, 8003)
    out = bench.cpu_baseline(d, 3.0, per_stratum=3)
    assert out["kind"] == "port" and out["cores"] >= 1 and out["extrapolated"] is True
    assert out["value"] > 0 and out["rows"] >= 2 and out["extrapolated_seconds_full_workload"] > 0
    # the sample covers several (type, k) strata and keeps the oracle's per-row results for the parity check
    typ, k, W = bench.row_strata(d["dat"])
    rows = out["row_index"]
    assert len({(int(typ[i]), int(k[i])) for i in rows}) >= 4
    ep, n_tot = d["eval_point"], 9
    th, dp, dm = ep[:81].reshape(9, 9), ep[81:90], ep[90:]
    i = int(rows[0])
    r = rr.patient_value_grad(th, dp, dm, d["dat"][i], want_grad=True)
    assert abs(r[1] - out["row_logp"][0]) <= 1e-14 * abs(r[1])
    assert out["row_grad_sum"].shape == (n_tot * (n_tot + 2),)
    # work estimate: a paired row with 3 + 2 events outweighs an unpaired row with 3
    assert W[(typ == 3)].max() > 0 and np.all(W >= 0)


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "6", "--patients", "300",
                          "--steps", "1", "--warmup", "1", "--ref-step-seconds", "3"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "patients/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["extrapolated"] is True
