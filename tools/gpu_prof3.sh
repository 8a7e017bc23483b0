#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ctc_" -s 3 -c 3 -f -o gpurun_out/prof_ctc python tools/run_path.py --ctc --B 64 --T 374 --U 80 --V 5000 --iters 2 > gpurun_out/ncu_ctc.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rnnt_alpha" -s 1 -c 1 -f -o gpurun_out/prof_lat python tools/run_path.py --iters 2 > gpurun_out/ncu_lat.log 2>&1; echo "rc=$?"
