#!/bin/bash
# fused CTC head: parity tests (bounded by timeout: a hung kernel is killed)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_sizes.py -m gpu -q -x -k "head" > gpurun_out/pytest_head.log 2>&1; echo "pytest head rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_head.log | tail -15
tail -50 gpurun_out/pytest_head.log | cut -c1-300
