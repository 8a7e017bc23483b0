#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1c}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"joint_|reduce_dpre" -s 4 -c 4 -f -o gpurun_out/prof_$TAG python tools/run_path.py --iters 2 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ctc.csv python tools/run_path.py --ctc --B 64 --T 374 --U 80 --V 5000 --iters 3 > /dev/null 2>&1; echo "ncu ctc rc=$?"
timeout 600 python bench.py --workload rnnt_cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rnnt_cfg4.json 2> gpurun_out/bench_rnnt_cfg4.err; echo "cfg4 rc=$?"; cat gpurun_out/bench_rnnt_cfg4.json; tail -3 gpurun_out/bench_rnnt_cfg4.err
timeout 600 python bench.py --lengths ragged --steps 10 --no-cpu-baseline > gpurun_out/bench_rnnt_cfg3_ragged.json 2>/dev/null; echo "ragged rc=$?"; cat gpurun_out/bench_rnnt_cfg3_ragged.json
