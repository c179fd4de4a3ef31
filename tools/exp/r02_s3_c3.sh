#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_c3.log; : > $L
for rep in 1 2; do
timeout 300 python tools/bench_extra.py c3 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('eval_ms',round(d['eval_common_kernel_ms'],4),'estep',{k:round(v,3) for k,v in d['e_step_phases_ms'].items() if v>0},'vem',{k:round(v,3) for k,v in d['vem_phases_ms'].items() if v>0})
" >> $L
done
timeout 900 python -m pytest tests -m gpu -x -q -k "edges or mstep or estep or baseline or fullsize" > gpurun_out/r02s3_c3_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02s3_c3_tests.log
