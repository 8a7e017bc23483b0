#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 600 python bench.py > gpurun_out/bench_${TAG}_rnnt_cfg3.json 2> gpurun_out/bench.err; echo "bench cfg3 rc=$?"; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --lengths ragged --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_rnnt_cfg3_ragged.json 2>/dev/null; echo "bench ragged rc=$?"
timeout 600 python bench.py --workload rnnt_cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_rnnt_cfg4.json 2>/dev/null; echo "bench cfg4 rc=$?"
for f in gpurun_out/bench_${TAG}_rnnt*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'roof', r.get('frac'), r.get('executed_frac'), 'bwd ms', r.get('kernel_ms'), 'fwd', (r.get('forward') or {}).get('kernel_ms'), (r.get('forward') or {}).get('frac'), 'step frac', d.get('step_frac_of_burst_peak'), d.get('clocks'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
