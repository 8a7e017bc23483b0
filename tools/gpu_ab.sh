#!/bin/bash
# A/B: previous library (libemoasr_b200_before.so) vs the working tree, alternating, fwd / bwd medians
ARGS=${AB_ARGS:-"--iters 12"}
for i in 1 2 3; do
EMOASR_B200_LIB=emoasr_b200/lib/libemoasr_b200_before.so timeout -k 10 300 python tools/time_joint.py $ARGS 2>&1 | grep "^joint" | sed 's/^/before /'
timeout -k 10 300 python tools/time_joint.py $ARGS 2>&1 | grep "^joint" | sed 's/^/after  /'
done
