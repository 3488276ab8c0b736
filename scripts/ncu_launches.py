"""Summarise an ncu --csv launch list (gpu__time_duration.sum, launch__grid_size): usage ncu_launches.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for i, r in enumerate(rows):
    if r and r[0] == 'ID': hdr = r; start = i + 1; break
idx = {h: i for i, h in enumerate(hdr)}
data = {}
for r in rows[start:]:
    if len(r) < len(hdr): continue
    data.setdefault(int(r[idx['ID']]), {'name': r[idx['Kernel Name']]})[r[idx['Metric Name']]] = float(r[idx['Metric Value']].replace(',', ''))
tot = {}
for k, v in sorted(data.items()):
    nm = v['name'].split('(')[0].replace('void ', '').replace('mmh::', '')
    if k < n: print(k, nm[:28], 'grid', int(v['launch__grid_size']), 'us %.1f' % (v['gpu__time_duration.sum'] / 1e3), 'ns/CTA %.0f' % (v['gpu__time_duration.sum'] / v['launch__grid_size']))
    tot[nm] = tot.get(nm, 0) + v['gpu__time_duration.sum'] / 1e6
print({k: round(v, 2) for k, v in tot.items()})
