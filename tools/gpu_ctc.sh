#!/bin/bash
# CTC path: parity tests, bench at cfg 2 / cfg 1, launch list (bounded by timeouts)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ctc or CTC" > gpurun_out/pytest_ctc.log 2>&1; echo "pytest ctc rc=$?"; tail -5 gpurun_out/pytest_ctc.log
timeout 300 python -m pytest tests -m gpu -x -q -k "dropin or golden or seam" > gpurun_out/pytest_ctc2.log 2>&1; echo "pytest other rc=$?"; tail -2 gpurun_out/pytest_ctc2.log
timeout 200 python bench.py --workload ctc_cfg2 --no-cpu-baseline > gpurun_out/bench_ctc2_quick.json 2>/dev/null; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_ctc2_quick.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ctc_quick.csv python tools/run_path.py --ctc --B 64 --T 374 --U 80 --V 5000 --iters 3 > /dev/null 2>&1; echo "ncu rc=$?"
