#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_gpu.log | tail -15
tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
