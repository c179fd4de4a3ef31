#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_det.log; : > $L
for i in 1 2 3 4 5 6; do timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $L; done
timeout 300 python tools/dbg/reuse_check.py >> $L 2>&1
