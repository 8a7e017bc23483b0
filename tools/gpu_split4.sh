#!/bin/bash
# role-split sweep of the ring backward with an un-instrumented tuning build (-DEMO_TUNING: EMO_RING_SPLIT honoured)
TUNE=emoasr_b200/lib/libemoasr_b200_tune.so
for p in 27 28 29 30 31 32 33 34; do
  d=$((50-p)); sp="$p,$d,6"
  echo -n "cfg3 split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_routes.py --iters 10 2>&1 | grep -E "^route" | cut -c30-110
done
for p in 22 24 25 26 27 28 29 30; do
  d=$((42-p)); sp="$p,$d,2"
  echo -n "cfg4 split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_routes.py --B 8 --T 1000 --U 400 --V 4096 --iters 3 2>&1 | grep -E "^route" | cut -c30-110
done
