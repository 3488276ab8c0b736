"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
    if r[ui] in ("ns", "nsecond"): v /= 1e6
    elif r[ui] in ("us", "usecond"): v /= 1e3
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.2f} ms {100*v[1]/tot:5.1f}%  {v[0]:6d} launches  {k}")
