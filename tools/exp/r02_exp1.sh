#!/bin/bash
# round 2, GPU call 1: sanity of the round-1 tree, the two untimed compile-time switches, ncu of the E-step kernels at C3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_exp1_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_exp1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_exp1_tests.log
L=magmaclust_b200/lib
for v in libmagmaclust_b200 variant_B variant_C; do
  echo "== $v overlapped" >> gpurun_out/r02_exp1_bench.log
  MCB_LIB=$PWD/$L/$v.so timeout 300 python tools/bench_phases.py --steps 20 --warmup 3 --no-steady >> gpurun_out/r02_exp1_bench.log 2>&1
  echo "== $v theta_i alone (MCB_NO_OVERLAP)" >> gpurun_out/r02_exp1_bench.log
  MCB_NO_OVERLAP=1 MCB_LIB=$PWD/$L/$v.so timeout 300 python tools/bench_phases.py --steps 20 --warmup 3 --no-steady >> gpurun_out/r02_exp1_bench.log 2>&1
done
timeout 600 python tools/bench_extra.py c3 > gpurun_out/r02_exp1_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"indiv_tau_kernel|indiv_factor2_kernel" -c 2 \
  -o gpurun_out/r02_estep_c3 python tools/bench_extra.py c3 > gpurun_out/r02_exp1_ncu.log 2>&1
echo done
