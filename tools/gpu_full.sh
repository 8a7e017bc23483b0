#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_gpu.log | tail -5
for mode in "" "--unfolded"; do
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 $mode > gpurun_out/bench_fold$mode.json 2>/dev/null; python - "$mode" <<PY
import json,sys
d=json.load(open('gpurun_out/bench_fold'+sys.argv[1]+'.json')); print('mode', sys.argv[1] or 'folded', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['step_frac_of_burst_peak'])
PY
done
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fold.csv python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_fold.csv | tail -32 | cut -c1-130
