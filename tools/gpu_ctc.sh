#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ctc or lattice or dense or known or aligner" > gpurun_out/pytest_ctc.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ctc.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ctc.csv python tools/run_path.py --ctc --B 64 --T 374 --U 80 --V 5000 --iters 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rnnt_ -c 12 --csv --log-file gpurun_out/launches_lat.csv python tools/run_path.py --iters 3 > /dev/null 2>&1
