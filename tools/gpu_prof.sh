#!/bin/bash
# ncu --set full capture of the three tcgen05 kernels (second iteration) + single-vs-pair forward timing
mkdir -p gpurun_out
TAG=${1:-r1a}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:joint_ -s 3 -c 3 -f -o gpurun_out/prof_$TAG python tools/run_path.py --iters 2 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
EMO_FWD_SINGLE_CTA=1 timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_single_$TAG.json 2>/dev/null; echo "single rc=$?"
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_pair_$TAG.json 2>/dev/null; echo "pair rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*_'+'*.json')):
    try:
        d=json.load(open(f)); print(f, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
    except Exception as e: print(f, e)
PY
ls -la gpurun_out
