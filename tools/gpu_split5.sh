#!/bin/bash
# role-split sweep after the ring-manager change (tuning build)
TUNE=emoasr_b200/lib/libemoasr_b200_tune.so
for sp in "25,25,6" "26,24,6" "27,23,6" "28,22,6" "29,21,6" "30,20,6" "24,22,7" "25,21,7" "26,20,7" "27,19,7" "29,25,5" "30,24,5" "31,23,5"; do
  echo -n "cfg3 split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_joint.py --iters 10 2>&1 | grep -E "^joint" | cut -c1-120
done
for sp in "20,22,2" "24,18,2" "28,14,2" "16,10,3" "20,6,3"; do
  echo -n "cfg4 split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_joint.py --B 8 --T 1000 --U 400 --V 4096 --iters 3 2>&1 | grep -E "^joint" | cut -c1-120
done
