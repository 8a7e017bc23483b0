#!/bin/bash
# issuer-wait instrumentation of the z-cache backward kernels (library built with -DEMO_ZC_PROF)
mkdir -p gpurun_out
timeout 120 python tools/run_path.py --iters 2 > gpurun_out/zcprof.log 2>&1; echo "rc=$?"; grep issuer gpurun_out/zcprof.log
