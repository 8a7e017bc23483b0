#!/bin/bash
# Evidence run: tests, smoke, bench lines, launch lists, ncu --set full captures.  Everything bounded by `timeout`.
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}_rnnt_cfg3.json 2> gpurun_out/bench.err; echo "bench cfg3 rc=$?"
timeout 600 python bench.py --lengths ragged --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_rnnt_cfg3_ragged.json 2>/dev/null; echo "bench ragged rc=$?"
timeout 600 python bench.py --workload rnnt_cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_rnnt_cfg4.json 2>/dev/null; echo "bench cfg4 rc=$?"
timeout 300 python bench.py --workload ctc_cfg2 > gpurun_out/bench_${TAG}_ctc_cfg2.json 2>/dev/null; echo "bench ctc2 rc=$?"
timeout 300 python bench.py --workload ctc_cfg1 > gpurun_out/bench_${TAG}_ctc_cfg1.json 2>/dev/null; echo "bench ctc1 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_rnnt_cfg3.json 2>/dev/null; echo "reference rc=$?"
timeout 300 python tools/time_decode.py > gpurun_out/${TAG}_time_decode.txt 2>&1; tail -1 gpurun_out/${TAG}_time_decode.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_rnnt_cfg3.csv python tools/run_path.py --iters 3 > /dev/null 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}_ctc_head_cfg2.csv python tools/run_path.py --ctc-head --B 64 --T 374 --U 80 --V 5000 --J 256 --iters 3 > /dev/null 2>&1; echo "ncu ctc head rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"joint_bwd_ring|joint_fwd|reduce_dh|rnnt_alpha" -s 4 -c 4 -f -o gpurun_out/prof_${TAG} python tools/run_path.py --iters 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"joint_bwd_ring|joint_fwd|head_|ctc_lattice" -s 8 -c 8 -f -o gpurun_out/prof_${TAG}_ctc_head python tools/run_path.py --ctc-head --B 64 --T 374 --U 80 --V 5000 --J 256 --iters 2 > /dev/null 2>&1; echo "ncu ctc head full rc=$?"
for f in gpurun_out/bench_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'roof', r.get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
