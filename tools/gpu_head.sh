#!/bin/bash
# fused CTC head: parity, timing, role-split sweep (instrumented library), launch list
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_sizes.py -m gpu -q -x -k "head" > gpurun_out/pytest_head.log 2>&1; echo "pytest head rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_head.log | tail -5
timeout -k 10 120 python tools/time_head.py
PROF=emoasr_b200/lib/libemoasr_b200_prof.so
if [ -f $PROF ]; then
for sp in ${SPLITS:-"40,14,1" "36,18,1" "44,10,1" "46,8,1" "30,24,1"}; do
  echo -n "split $sp: "
  EMO_RING_SPLIT=$sp EMOASR_B200_LIB=$PROF timeout -k 10 120 python tools/time_head.py --iters 6 2>&1 | grep -E "^head"
done
fi
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_ctc_head.csv python tools/run_path.py --ctc-head --B 64 --T 374 --U 80 --V 5000 --J 256 --iters 3 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_ctc_head.csv | grep emo
