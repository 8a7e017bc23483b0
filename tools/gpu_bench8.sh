#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
TAG=${2:-r2d}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_rnnt_cfg3_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench $N rc=$?"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --grad-payload-mb 99.4 > gpurun_out/bench_${TAG}_rnnt_cfg3_${N}gpu_payload103.json 2>> gpurun_out/bench_${N}gpu.err; echo "bench $N payload rc=$?"
timeout -k 10 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/bench_${TAG}_rnnt_cfg3_1gpu_same_box8.json 2>/dev/null
python - <<PY
import json
for f in ("bench_${TAG}_rnnt_cfg3_${N}gpu", "bench_${TAG}_rnnt_cfg3_${N}gpu_payload103", "bench_${TAG}_rnnt_cfg3_1gpu_same_box8"):
    try:
        d=json.load(open('gpurun_out/'+f+'.json'))
        print(f, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], (d.get('extra') or {}).get('per_rank_local_step_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
