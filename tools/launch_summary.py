"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, mean, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
c = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    k = r[c['Kernel Name']][:72]
    v = float(r[c['Metric Value']].replace(',', ''))
    u = r[c['Metric Unit']]
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v):9.1f}us  share={100*sum(v)/tot:5.1f}%")
