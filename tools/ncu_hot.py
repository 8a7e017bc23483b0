"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv): top SASS lines by
stall samples with their dominant stall reason, plus totals per stall reason."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = {s: 0 for s in stalls}
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    n = int(r[col["# Samples"]] or 0)
    ex = int(r[col["Instructions Executed"]] or 0)
    st = {s: int(r[col[s]] or 0) for s in stalls}
    for s in stalls:
        tot[s] += st[s]
    data.append((n, ex, r[col["Source"]].strip(), st, len(data)))
total = sum(d[0] for d in data)
print("total samples", total)
print("by reason:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
top = sorted(data, key=lambda d: -d[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for n, ex, src, st, idx in top:
    dom = max(st.items(), key=lambda kv: kv[1])
    print(f"{idx:5d} {n:7d} {100.0*n/total:5.1f}%  ex={ex:9d}  {dom[0]}={dom[1]:6d}  {src[:90]}")
