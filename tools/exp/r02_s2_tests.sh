#!/bin/bash
# round 2, session 2, call 1: GPU tests of the new paths (big individuals, k-means, R shim chain), bench sanity, ncu of tau2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > gpurun_out/r02s2_stub_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02s2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r02s2_bench.json 2> gpurun_out/r02s2_bench.err
echo "rc=$?" >> gpurun_out/r02s2_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"indiv_tau2_kernel" -c 1 \
  -o gpurun_out/r02s2_tau2_c3 python tools/bench_extra.py c3 > gpurun_out/r02s2_ncu_tau2.log 2>&1
echo done
