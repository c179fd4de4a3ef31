#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_exp2.log; : > $L
for rep in 1 2; do for v in 3 4; do
echo "== v$v c3 (rep $rep)" >> $L
MCB_LIB=$PWD/build/variants/libmcb_v$v.so timeout 300 python tools/bench_extra.py c3 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('eval_ms',round(d['eval_common_kernel_ms'],4),'estep',{k:round(v,3) for k,v in d['e_step_phases_ms'].items() if v>0},'vem',{k:round(v,3) for k,v in d['vem_phases_ms'].items() if v>0})
" >> $L
done; done
for v in 4; do echo "== v$v bench C2" >> $L
MCB_LIB=$PWD/build/variants/libmcb_v$v.so timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-steady --skip-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'])
" >> $L
done
