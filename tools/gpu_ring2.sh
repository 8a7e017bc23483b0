#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16" > gpurun_out/pytest_ring.log 2>&1; echo "pytest bf16 rc=$?"
tail -4 gpurun_out/pytest_ring.log
timeout -k 10 200 python tools/time_routes.py --iters 8 2>&1 | grep -E "^route|diff"
PROF=emoasr_b200/lib/libemoasr_b200_prof.so
EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_routes.py --routes ring --iters 2 2>&1 | grep -E "^ring" | sort | uniq | tail -12
for sp in "24,26,6" "25,25,6" "26,24,6" "27,23,6" "28,22,6" "29,21,6" "24,22,7" "25,21,7" "23,23,7"; do
  echo -n "split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_routes.py --routes ring --iters 6 2>&1 | grep -E "^route" 
done
timeout -k 10 900 python -m pytest tests/test_gpu_sizes.py -m gpu -x -q > gpurun_out/pytest_sizes.log 2>&1; echo "pytest sizes rc=$?"
tail -15 gpurun_out/pytest_sizes.log
