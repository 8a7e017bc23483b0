#!/bin/bash
# first GPU check of the ring backward (bounded: a hung kernel is killed by timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16_loss_and_grads and ring" > gpurun_out/pytest_ring.log 2>&1; echo "pytest ring rc=$?"
tail -25 gpurun_out/pytest_ring.log
timeout -k 10 200 python tools/time_routes.py > gpurun_out/time_routes.log 2>&1; echo "time_routes rc=$?"; tail -5 gpurun_out/time_routes.log
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -8 gpurun_out/pytest_all.log
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_ring.csv python tools/run_path.py --iters 2 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_ring.csv | tail -20
