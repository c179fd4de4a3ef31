#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_estep.py tests/test_gpu_edges.py tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r02_tau_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02_tau_tests.log
timeout 300 python tools/bench_extra.py c3 > gpurun_out/r02_tau_c3.log 2>&1
MCB_TAU_V1=1 timeout 300 python tools/bench_extra.py c3 > gpurun_out/r02_tau_c3_v1.log 2>&1
timeout 300 python tools/bench_phases.py --steps 20 --warmup 3 --no-steady > gpurun_out/r02_tau_c2.log 2>&1
