#!/bin/bash
# ncu --set full capture of the tcgen05 kernels (second iteration of tools/run_path.py)
mkdir -p gpurun_out
TAG=${1:-r1}
REGEX=${2:-joint_}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 3 -c 3 -f -o gpurun_out/prof_$TAG python tools/run_path.py --iters 2 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
