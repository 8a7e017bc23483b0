#!/bin/bash
# role wait accounting + role-split sweep of the ring backward (instrumented library), then ncu --set full
mkdir -p gpurun_out
PROF=emoasr_b200/lib/libemoasr_b200_prof.so
echo "== wait accounting (default split)"
EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_joint.py --iters 2 2>&1 | grep -E "ring|joint" | tail -12
for sp in "26,24,6" "28,22,6" "29,21,6" "30,20,6" "31,19,6" "32,22,5" "34,20,5" "30,24,5" "27,19,7"; do
  echo "== split $sp"
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_joint.py --iters 6 2>&1 | grep -E "^joint" 
done
echo "== ncu full (product library)"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"joint_bwd_ring|joint_fwd|reduce_dh" -s 3 -c 3 -f -o gpurun_out/prof_r2a python tools/run_path.py --iters 2 > gpurun_out/ncu_full_r2a.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_r2a.ncu-rep
