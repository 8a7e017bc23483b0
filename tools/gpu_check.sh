#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Everything is bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_rnnt_cfg3.json 2> gpurun_out/bench_rnnt_cfg3.err; echo "bench rc=$?"
timeout 300 python bench.py --workload ctc_cfg2 > gpurun_out/bench_ctc_cfg2.json 2> gpurun_out/bench_ctc_cfg2.err; echo "bench ctc rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_rnnt.csv python tools/run_path.py --iters 3 > gpurun_out/ncu_rnnt.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -5; cat gpurun_out/bench_rnnt_cfg3.json; tail -3 gpurun_out/bench_rnnt_cfg3.err; cat gpurun_out/bench_ctc_cfg2.json
