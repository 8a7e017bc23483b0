#!/bin/bash
# parity of the bf16 joint + quick bench + launch list (bounded by timeout)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -m gpu -q -x -k "bf16 or cfg3 or cfg4 or from_outputs or golden or lse_output" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_quick.log | tail -5
for i in 1 2; do
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 30 > gpurun_out/bench_quick.json 2>/dev/null; python - <<PY
import json
d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']; print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'bwd', r['kernel_ms'], 'fwd', r['forward']['kernel_ms'], d['step_frac_of_burst_peak'], d['clocks']['reasons'])
PY
done
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_quick.csv python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_quick.csv | grep "emo" | cut -c1-130
