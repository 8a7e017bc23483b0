#!/bin/bash
# issuer-wait instrumentation of the tcgen05 kernels (library built with EMO_NVCC_EXTRA=-DEMO_ZC_PROF);
# EMO_ZC_DEBUG switches (backward kernels): 1 = no transform math, 2 = no MMAs, 3 = neither (pure data movement),
# 4 = dWz reads its A operand K-major (wrong math; isolates the cost of the MN-major A operand)
mkdir -p gpurun_out
for f in ${ZC_FLAGS:-0}; do
  echo "== EMO_ZC_DEBUG=$f"
  EMO_ZC_DEBUG=$f timeout 120 python tools/run_path.py --iters 2 > gpurun_out/zcprof_$f.log 2>&1; echo "rc=$?"; grep issuer gpurun_out/zcprof_$f.log | tail -3
done
