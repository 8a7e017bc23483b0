#!/bin/bash
# Evidence of the DEFAULT bench command (folded path): bench line, its ncu launch list and one ncu --set full capture.
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 600 python bench.py > gpurun_out/bench_${TAG}_rnnt_cfg3.json 2> gpurun_out/bench.err; echo "bench cfg3 rc=$?"
timeout 600 python bench.py --workload rnnt_cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_rnnt_cfg4.json 2>/dev/null; echo "bench cfg4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}_bench_rnnt_cfg3.csv python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"joint_bwd_ring|joint_fwd|reduce_dh|rnnt_alpha|proj_gemm|cast_colsum|multi_cast" -s 33 -c 11 -f -o gpurun_out/prof_${TAG}_bench python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
for f in ("bench_${TAG}_rnnt_cfg3", "bench_${TAG}_rnnt_cfg4"):
    d=json.load(open('gpurun_out/'+f+'.json')); r=d['roofline']
    print(f, d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('regions_utt_s_rank0'), 'bwd', r['kernel_ms'], r['frac'], r['executed_frac'], 'fwd', r['forward']['kernel_ms'], r['forward']['frac'], 'step frac', d['step_frac_of_burst_peak'], d['extra'].get('step_ms_rank0'))
PY
