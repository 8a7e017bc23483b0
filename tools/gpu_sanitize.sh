#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small fwd+bwd of both paths
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/run_path.py --B 2 --T 20 --U 9 --V 288 --J 128 --iters 1 > gpurun_out/sanitizer_${tool}_rnnt.log 2>&1; echo "$tool rnnt rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_rnnt.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/run_path.py --ctc --B 3 --T 30 --U 8 --V 100 --iters 1 > gpurun_out/sanitizer_${tool}_ctc.log 2>&1; echo "$tool ctc rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_ctc.log
done
