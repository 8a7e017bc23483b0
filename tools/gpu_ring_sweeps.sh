#!/bin/bash
# role-split and slot-count sweeps of the ring backward with an un-instrumented tuning build
# (python -c "from emoasr_b200 import build; build.build_library(force=True, out='emoasr_b200/lib/libemoasr_b200_tune.so', extra=['-DEMO_TUNING'])":
#  EMO_RING_SPLIT / EMO_RING_SLOTS are honoured by tuning builds only)
TUNE=emoasr_b200/lib/libemoasr_b200_tune.so
for p in 27 28 29 30 31 32 33 34; do
  d=$((50-p)); sp="$p,$d,6"
  echo -n "cfg3 split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_joint.py --iters 10 2>&1 | grep -E "^joint" | cut -c1-120
done
for p in 22 24 25 26 27 28 29 30; do
  d=$((42-p)); sp="$p,$d,2"
  echo -n "cfg4 split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_joint.py --B 8 --T 1000 --U 400 --V 4096 --iters 3 2>&1 | grep -E "^joint" | cut -c1-120
done
# ring slots (EMO_RING_SLOTS)
for n in 32 40 48 64 80 96 128; do
  echo -n "cfg3 slots $n: "
  EMO_RING_SLOTS=$n EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_joint.py --iters 10 2>&1 | grep -E "^joint" | cut -c1-120
done
for n in 32 48 64 96 128; do
  echo -n "cfg4 slots $n: "
  EMO_RING_SLOTS=$n EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_joint.py --B 8 --T 1000 --U 400 --V 4096 --iters 3 2>&1 | grep -E "^joint" | cut -c1-120
done
