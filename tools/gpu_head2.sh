#!/bin/bash
mkdir -p gpurun_out
echo "== cfg4 before / after"
EMOASR_B200_LIB=emoasr_b200/lib/libemoasr_b200_before.so timeout -k 10 200 python tools/time_routes.py --B 8 --T 1000 --U 400 --V 4096 --iters 4 2>&1 | grep "^route"
timeout -k 10 200 python tools/time_routes.py --B 8 --T 1000 --U 400 --V 4096 --iters 4 2>&1 | grep "^route"
echo "== cfg3 before / after"
EMOASR_B200_LIB=emoasr_b200/lib/libemoasr_b200_before.so timeout -k 10 200 python tools/time_routes.py --iters 8 2>&1 | grep "^route"
timeout -k 10 200 python tools/time_routes.py --iters 8 2>&1 | grep "^route"
echo "== head"
timeout -k 10 120 python tools/time_head.py
PROF=emoasr_b200/lib/libemoasr_b200_prof.so
for sp in "40,14,1" "36,18,1" "32,22,1" "28,26,1" "44,10,1"; do
  echo -n "split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_head.py --iters 6 2>&1 | grep -E "^head"
done
timeout -k 10 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_gpu.log | tail -5
