#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/dbg/reuse_check.py > gpurun_out/r02s2_reuse.log 2>&1
MCB_NO_OVERLAP=1 timeout 300 python tools/dbg/reuse_check.py > gpurun_out/r02s2_reuse_noov.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_baseline_configs.py::test_training_then_posterior_on_a_larger_grid_then_training_again > gpurun_out/r02s2_tests2.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_tests2.log
