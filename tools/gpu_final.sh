#!/bin/bash
# End-of-round evidence run: tests, smoke, every bench line, launch lists, one ncu --set full capture.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/bench_${TAG}_rnnt_cfg3.json 2> gpurun_out/bench.err; echo "bench cfg3 rc=$?"
timeout 600 python bench.py --lengths ragged --no-cpu-baseline > gpurun_out/bench_${TAG}_rnnt_cfg3_ragged.json 2>/dev/null; echo "bench ragged rc=$?"
timeout 600 python bench.py --workload rnnt_cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_rnnt_cfg4.json 2>/dev/null; echo "bench cfg4 rc=$?"
timeout 300 python bench.py --workload ctc_cfg2 > gpurun_out/bench_${TAG}_ctc_cfg2.json 2>/dev/null; echo "bench ctc2 rc=$?"
timeout 300 python bench.py --workload ctc_cfg1 > gpurun_out/bench_${TAG}_ctc_cfg1.json 2>/dev/null; echo "bench ctc1 rc=$?"
timeout 300 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_rnnt_cfg3_fp32.json 2>/dev/null; echo "bench fp32 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_rnnt_cfg3.json 2>/dev/null; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_rnnt_cfg3.csv python tools/run_path.py --iters 3 > /dev/null 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}_ctc_cfg2.csv python tools/run_path.py --ctc --B 64 --T 374 --U 80 --V 5000 --iters 3 > /dev/null 2>&1; echo "ncu ctc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"joint_|reduce_d|rnnt_alpha" -s 5 -c 5 -f -o gpurun_out/prof_${TAG} python tools/run_path.py --iters 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"ctc_" -s 3 -c 3 -f -o gpurun_out/prof_${TAG}_ctc python tools/run_path.py --ctc --B 64 --T 374 --U 80 --V 5000 --iters 2 > /dev/null 2>&1; echo "ncu ctc full rc=$?"
for f in gpurun_out/bench_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'roof', r.get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
