#!/bin/bash
TUNE=emoasr_b200/lib/libemoasr_b200_tune.so
for n in 32 40 48 64 80 96 128; do
  echo -n "cfg3 slots $n: "
  EMO_RING_SLOTS=$n EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_routes.py --iters 10 2>&1 | grep -E "^route" | cut -c30-110
done
for n in 32 48 64 96 128; do
  echo -n "cfg4 slots $n: "
  EMO_RING_SLOTS=$n EMOASR_B200_LIB=$TUNE timeout -k 10 200 python tools/time_routes.py --B 8 --T 1000 --U 400 --V 4096 --iters 3 2>&1 | grep -E "^route" | cut -c30-110
done
