#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_sizes.py -m gpu -q -x -k "step or greedy or align or distill or lse_output" > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_dec.log | tail -5
tail -40 gpurun_out/pytest_dec.log | cut -c1-250
