#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s2_tma2.log
: > $L
timeout 400 tools/dbg/tma_check 4096 0 3000 0 >> $L 2>&1
timeout 400 tools/dbg/tma_check 4096 0 2000 1 >> $L 2>&1
timeout 400 tools/dbg/tma_check 8192 0 600 0 >> $L 2>&1
echo "== dense_check (TMA)" >> $L; timeout 300 tools/dbg/dense_check 2>&1 | grep -E "spd_inverse N=(800|1000|1111)|CUDA" >> $L
for n in 4096 8192; do
  for r in 1 2 3; do echo "== dense_time $n (TMA) run $r" >> $L; timeout 300 tools/dbg/dense_time $n 1 2>&1 | grep -E "build \+ inverse|info|CUDA" >> $L; done
  echo "== dense_time $n (MCB_NO_TMA)" >> $L; MCB_NO_TMA=1 timeout 300 tools/dbg/dense_time $n 1 2>&1 | grep -E "build \+ inverse|info|CUDA" >> $L
done
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_estep.py tests/test_gpu_baseline_configs.py -m gpu -x -q -k "c4 or large_grid" >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 300 python tools/bench_extra.py c4 >> $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_update_tma_kernel" -c 1 -s 4 \
  -o gpurun_out/r02s2_tma_update2 tools/dbg/dense_time 8192 1 > gpurun_out/r02s2_ncu_tma2.log 2>&1
echo done >> $L
