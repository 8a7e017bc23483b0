#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','step_frac_of_burst_peak','clocks')})
r=d['roofline']; print({k:r[k] for k in ('kernel_ms','frac','executed_frac','traffic')}, r['forward']['kernel_ms'], r['forward']['frac'])
print(d.get('extra')); print(d.get('gpu_baseline')); print(d.get('cpu_baseline'))
PY
