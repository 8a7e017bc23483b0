"""DRAM traffic per kernel from `ncu -i X.ncu-rep --page raw --csv` -> one entry of profiles/dram_traffic.json
(the file bench.py reads `roofline.traffic` from).

    python tools/ncu_traffic.py <raw.csv> <workload> <route> <source-note> [--launches-per-step N]

Kernels are attributed to the forward call (emo_rnnt_joint_fwd + lattice) or the backward call
(emo_rnnt_joint_bwd) by name; the capture must hold whole steps (N launches of every kernel per step are averaged).
"""
import csv
import json
import os
import sys

raw, workload, route, note = sys.argv[1:5]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
c = {h: i for i, h in enumerate(hdr)}
unit = {h: rows[1][i] for i, h in enumerate(hdr)}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


per = {}
for r in rows[2:]:
    name = r[c["Kernel Name"]]
    short = name.split("(")[0].split("::")[-1].replace("void ", "")
    b = to_bytes(r[c["dram__bytes_read.sum"]], unit["dram__bytes_read.sum"]) + \
        to_bytes(r[c["dram__bytes_write.sum"]], unit["dram__bytes_write.sum"])
    per.setdefault(short, []).append(b)
kern = {k: sum(v) / len(v) for k, v in per.items()}
fwd_names = ("joint_fwd_kernel", "rnnt_alpha_beta_kernel", "rnnt_gamma_kernel", "ctc_row", "ctc_lattice",
             "multi_cast_kernel", "proj_gemm_kernel<0, 0>")   # folded path: the call's casts and the forward projections
fwd = sum(v for k, v in kern.items() if k.startswith(fwd_names))
bwd = sum(v for k, v in kern.items() if not k.startswith(fwd_names))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "dram_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d.setdefault(workload, {})[route] = {"fwd": fwd, "bwd": bwd, "step": fwd + bwd,
                                     "kernels": {k: round(v) for k, v in kern.items()}, "capture": note}
d["_source"] = "ncu --set full captures under profiles/ (tools/ncu_traffic.py); per-entry 'capture' names the file"
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(d[workload][route], indent=1))
