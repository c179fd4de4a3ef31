#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"indiv_eval2_kernel" -c 1 -s 0 -f \
  -o gpurun_out/r02s3_eval2_c3 python tools/bench_extra.py c3 > gpurun_out/r02s3_ncu_eval2.log 2>&1
echo done
