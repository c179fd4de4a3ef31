#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s2_panel.log
: > $L
echo "== dense_check (new panel)" >> $L; timeout 300 tools/dbg/dense_check 2>&1 | grep -E "spd_inverse N=(800|1000|1111)|CUDA" >> $L
echo "== dense_check (MCB_OLD_PANEL)" >> $L; MCB_OLD_PANEL=1 timeout 300 tools/dbg/dense_check 2>&1 | grep -E "spd_inverse N=(800|1000|1111)|CUDA" >> $L
for n in 2000 4096 8192; do
  for v in "MCB_NOP=1" "MCB_OLD_PANEL=1"; do echo "== dense_time $n $v" >> $L; env $v timeout 300 tools/dbg/dense_time $n 1 2>&1 | grep -E "build \+ inverse|info|CUDA" >> $L; done
done
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_estep.py tests/test_gpu_baseline_configs.py -m gpu -x -q -k "c4 or large_grid" >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 300 python tools/bench_extra.py c4 >> $L 2>&1
echo done >> $L
