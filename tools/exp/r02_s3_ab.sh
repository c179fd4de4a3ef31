#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_ab.log; : > $L
for rep in 1 2; do for b in dense_time dense_time_oldp; do for z in 1 5; do
echo "== $b 401 $z" >> $L; timeout 120 tools/dbg/$b 401 $z 2>&1 | grep "build + inverse" >> $L; done; done; done
echo "== slab_prof" >> $L; timeout 120 tools/dbg/slab_prof >> $L 2>&1
