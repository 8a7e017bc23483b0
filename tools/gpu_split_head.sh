#!/bin/bash
# role-split sweep of the ring kernel in the CTC head's plain mode (cfg 2: J = 256, V = 5000 -> 20 vocab roles, 5 groups)
TUNE=emoasr_b200/lib/libemoasr_b200_tune.so
echo -n "head default: "; EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_head.py --iters 10 2>&1 | grep -E "^head"
for sp in "30,24,1" "35,19,1" "40,14,1" "45,9,1" "50,4,1" "25,9,2" "20,14,2" "15,19,2"; do
  echo -n "head split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_head.py --iters 10 2>&1 | grep -E "^head"
done
