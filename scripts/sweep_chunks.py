"""Sweep chunk budget x side streams for one workload (each combination in a fresh process, MMH_STREAMS is read at create)."""
import sys, os, subprocess
n, nd = sys.argv[1], sys.argv[2]
for chunk in sys.argv[3].split(","):
    for ns in sys.argv[4].split(","):
        env = dict(os.environ, MMH_STREAMS=ns, EVALS="4", VALUE_ONLY="0")
        out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "quick_time.py"), n, nd, chunk],
                             env=env, capture_output=True, text=True).stdout
        ms = [float(l.split("dev_ms")[1].split()[0]) for l in out.splitlines() if l.startswith("eval")]
        print(f"chunk={float(chunk)/2**20:.0f}MiB streams={ns}: dev_ms {min(ms[1:]) if len(ms) > 1 else ms}", flush=True)
