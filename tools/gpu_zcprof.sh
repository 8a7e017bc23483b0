#!/bin/bash
# issuer-wait instrumentation of the z-cache backward kernels (library built with -DEMO_ZC_PROF);
# EMO_ZC_DEBUG switches: 1 = no transform math, 2 = no MMAs, 3 = neither (pure data movement)
mkdir -p gpurun_out
for f in 0 1 2 3; do
  echo "== EMO_ZC_DEBUG=$f"
  EMO_ZC_DEBUG=$f timeout 120 python tools/run_path.py --iters 2 > gpurun_out/zcprof_$f.log 2>&1; echo "rc=$?"; grep issuer gpurun_out/zcprof_$f.log | tail -2
done
