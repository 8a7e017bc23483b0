#!/bin/bash
# quick GPU check of the bf16 joint kernels (bounded: a hung kernel is killed by timeout)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16" > gpurun_out/pytest_bf16.log 2>&1; echo "pytest bf16 rc=$?"
tail -15 gpurun_out/pytest_bf16.log
timeout 120 python tools/run_path.py --iters 2 > gpurun_out/run_path.log 2>&1; echo "run_path rc=$?"; tail -2 gpurun_out/run_path.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cat gpurun_out/bench_quick.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_quick.csv python tools/run_path.py --iters 2 > /dev/null 2>&1; echo "ncu rc=$?"
