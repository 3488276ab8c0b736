"""Host emulation of the blocked substitution solve: the per-lane functions of metmhn_b200/csrc/mmh_blk.cuh compiled
with g++ and run lane by lane against a sequential substitution (tests/host/blk_host_test.cpp).  No GPU needed: this is
the check of the kernel's index arithmetic, skew schedule and rate bookkeeping that runs on every CPU test pass."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_solver_emulation_matches_sequential_substitution(tmp_path):
    exe = str(tmp_path / "blk_host_test")
    src = os.path.join(ROOT, "tests", "host", "blk_host_test.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert out.stdout.strip().endswith("OK")
