#!/bin/bash
PROF=emoasr_b200/lib/libemoasr_b200_prof.so
EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_routes.py --routes ring --iters 2 2>&1 | grep -E "^ring|^fwd|^route" | sort | uniq | grep -v "W dz loader" | tail -24
