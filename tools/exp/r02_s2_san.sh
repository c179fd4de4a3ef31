#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > /dev/null 2>&1
MCB_TMA_MIN_N=0 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/dbg/sanitize_r02.py > gpurun_out/r02s2_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_memcheck.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s2_tests6.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_tests6.log
timeout 900 python bench.py > gpurun_out/r02s2_bench_final1.json 2> gpurun_out/r02s2_bench_final1.err
echo "rc=$?" >> gpurun_out/r02s2_bench_final1.err
