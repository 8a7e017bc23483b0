"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv [--kernel-name ...]): per
kernel section, the top SASS lines by stall samples with their dominant stall reason, plus totals
per stall reason."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, hdr_i in enumerate(starts):
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = {s: 0 for s in stalls}
    for r in rows[hdr_i + 1:end]:
        if len(r) < len(hdr) or not r[col["# Samples"]].strip().isdigit():
            continue
        n = int(r[col["# Samples"]] or 0)
        ex = int(r[col["Instructions Executed"]] or 0)
        st = {s: int(r[col[s]] or 0) for s in stalls}
        for s in stalls:
            tot[s] += st[s]
        data.append((n, ex, r[col["Source"]].strip(), st, len(data)))
    total = sum(d[0] for d in data) or 1
    print(f"== section {si}: total samples {total}")
    print("by reason:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
    for n, ex, src, st, idx in sorted(data, key=lambda d: -d[0])[:topn]:
        dom = max(st.items(), key=lambda kv: kv[1])
        print(f"{idx:5d} {n:7d} {100.0*n/total:5.1f}%  ex={ex:9d}  {dom[0]}={dom[1]:6d}  {src[:90]}")
