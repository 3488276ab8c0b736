"""Aggregate the source page of an .ncu-rep (cuda,sass view) per CUDA source line: stall samples and instructions.
usage: ncu_lines.py file.ncu-rep [launch-index] [top-n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "0"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}
fname = None
hdr = None
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; idx = {h: i for i, h in enumerate(hdr)}; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":
        cur = (fname or "?", int(r[0]), r[1].strip())
        agg.setdefault(cur, [0.0, 0.0, {}])
        continue
    if cur is None or r[2] in ("...", ""):
        continue
    try:
        s = float(r[iS]); n = float(r[iI])
    except ValueError:
        continue
    a = agg[cur]
    a[0] += s; a[1] += n
    for i in stall_cols:
        try:
            v = float(r[i])
        except (ValueError, IndexError):
            continue
        if v: a[2][hdr[i]] = a[2].get(hdr[i], 0.0) + v
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print(f"total samples {ts:.0f}  warp instructions {ti:.0f}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
    print(f"{k[0][:18]:18s}:{k[1]:4d} {100 * a[0] / ts:5.1f}% smp {100 * a[1] / ti:5.1f}% ins  {' '.join(f'{n[6:]}={v:.0f}' for n, v in st):40s} | {k[2][:90]}")
