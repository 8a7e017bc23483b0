#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench $N rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','n_gpus')}, d['config']['parallelism'])
PY
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --grad-payload-mb 0 > gpurun_out/bench_${N}gpu_nopayload.json 2>> gpurun_out/bench_${N}gpu.err; echo "bench $N nopayload rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu_nopayload.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','n_gpus')})
PY
timeout -k 10 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/bench_1gpu_same_box.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_same_box.json')); print('1 GPU same box', d['value'], d['ms_per_step'], d['e2e']['value'])"
