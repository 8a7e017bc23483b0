"""Warp-stall samples per SOURCE LINE of one kernel: joins `ncu --page source --csv` (SASS rows with sampling
counts, in instruction order) with `nvdisasm -g` line info of the same cubin.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > k.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g -c file.cubin > k.sass
    python tools/ncu_lines.py k.csv k.sass <kernel-name-substring> [top]
"""
import collections
import csv
import re
import sys

csv_path, sass_path, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csv_path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
c = {h: i for i, h in enumerate(hdr)}
samp = [(r[c["Source"]], int(r[c["# Samples"]] or 0), int(r[c["Instructions Executed"]] or 0)) for r in rows[hi + 1:] if len(r) == len(hdr)]
# nvdisasm: walk the function, remember the current "//## File ..., line N" annotation (innermost inline frame last)
lines = open(sass_path).read().split("\n")
cur_file, cur_line, in_fn, seq = None, None, False, []
for ln in lines:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        in_fn = kname in ln
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        seq.append((cur_file, cur_line))
print(f"{len(samp)} sampled SASS rows, {len(seq)} disassembled instructions")
n = min(len(samp), len(seq))
agg = collections.Counter()
execd = collections.Counter()
for (src, s, ex), (f, l) in zip(samp[:n], seq[:n]):
    agg[(f, l)] += s
    execd[(f, l)] += ex
tot = sum(agg.values())
for (f, l), s in agg.most_common(top):
    print(f"{f}:{l:<5d} samples {s:8d} ({100.0 * s / tot:5.1f}%)  inst {execd[(f, l)]}")
