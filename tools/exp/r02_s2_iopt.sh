#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s2_tests9.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_tests9.log
L=gpurun_out/r02s2_iopt.log; : > $L
for v in "MCB_NOP=1" "MCB_HOST_THETA_I=1"; do echo "== $v" >> $L; env $v timeout 300 python tools/bench_extra.py c3 >> $L 2>&1; done
for v in "MCB_NOP=1" "MCB_HOST_THETA_I=1"; do echo "== C1 $v" >> $L; env $v timeout 300 python bench.py --workload C1 --steps 20 --warmup 3 --no-extras --skip-cpu-baseline --no-steady 2>&1 | tail -1 | cut -c1-300 >> $L; done
