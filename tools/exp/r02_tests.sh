#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q --durations=10 > gpurun_out/r02_tests_baseline.log 2>&1
echo "rc=$?" >> gpurun_out/r02_tests_baseline.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_baseline_configs.py > gpurun_out/r02_tests_all.log 2>&1
echo "rc=$?" >> gpurun_out/r02_tests_all.log
