#!/bin/bash
# fused CTC head: parity, timing, bench lines, launch list
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_sizes.py -m gpu -q -x -k "head" > gpurun_out/pytest_head.log 2>&1; echo "pytest head rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_head.log | tail -5
timeout -k 10 120 python tools/time_head.py
for wl in ctc_cfg2 ctc_cfg1; do
timeout -k 10 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$wl.json')); r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','dtype')}); print({k:r.get(k) for k in ('bound','achieved','frac','head','unfused_ms_per_step','hbm_equivalent')})
PY
done
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_ctc_head.csv python tools/run_path.py --ctc-head --B 64 --T 374 --U 80 --V 5000 --J 256 --iters 3 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_ctc_head.csv | grep emo
