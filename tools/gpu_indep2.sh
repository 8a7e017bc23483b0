#!/bin/bash
# two INDEPENDENT single-GPU bench processes at the same time (no torch.distributed): is the per-GPU slowdown of the
# multi-GPU runs there without any collective / process group?
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 60 > gpurun_out/indep_a.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 60 > gpurun_out/indep_b.json 2>/dev/null &
wait
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 60 > gpurun_out/indep_a_alone.json 2>/dev/null
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 60 > gpurun_out/indep_b_alone.json 2>/dev/null
python - <<PY
import json
for f in ("indep_a","indep_b","indep_a_alone","indep_b_alone"):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['value'], d['clocks'], 'bwd', d['roofline']['kernel_ms'], 'fwd', d['roofline']['forward']['kernel_ms'])
PY
nvidia-smi --query-gpu=index,power.limit,power.max_limit,enforced.power.limit --format=csv
