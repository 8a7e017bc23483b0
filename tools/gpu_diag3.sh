#!/bin/bash
python tools/diag_nccl.py none 2>&1 | grep "^rank"
python tools/diag_nccl.py nccl 2>&1 | grep "^rank"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/diag_nccl.py nccl 2>&1 | grep "^rank"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/diag_nccl.py gloo 2>&1 | grep "^rank"
