#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest parity rc=$?"
tail -4 gpurun_out/pytest_all.log
timeout -k 10 200 python tools/time_routes.py --iters 10 2>&1 | grep -E "^route|diff"
PROF=emoasr_b200/lib/libemoasr_b200_prof.so
EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_routes.py --routes ring --iters 2 2>&1 | grep -E "^ring|^fwd" | sort | uniq | grep -v "W dz loader" | tail -16
timeout -k 10 900 python -m pytest tests/test_gpu_sizes.py -m gpu -q > gpurun_out/pytest_sizes.log 2>&1; echo "pytest sizes rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_sizes.log | tail -15
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_ring.csv python tools/run_path.py --iters 2 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_ring.csv | grep emo
