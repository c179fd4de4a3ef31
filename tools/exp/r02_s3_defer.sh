#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_defer.log; : > $L
timeout 120 tools/dbg/dense_check > gpurun_out/r02s3_dc.log 2>&1; rc=$?
grep "spd_inverse" gpurun_out/r02s3_dc.log | awk '{print $2, $3, $6, $NF}' | tr '\n' ';' >> $L; echo >> $L; echo "dense_check rc=$rc" >> $L
if [ $rc -ne 0 ]; then exit 1; fi
for rep in 1 2; do for nd in 0 1; do for z in 1 5; do
if [ $nd = 1 ]; then export MCB_SLAB_NO_DEFER=1; else unset MCB_SLAB_NO_DEFER; fi
echo -n "no_defer=$nd Z=$z: " >> $L; timeout 120 tools/dbg/dense_time 401 $z 2>&1 | grep "build + inverse" >> $L
done; done; done
unset MCB_SLAB_NO_DEFER
for n in 200 600 768; do echo -n "N=$n defer: " >> $L; timeout 120 tools/dbg/dense_time $n 1 2>&1 | grep "build + inverse" >> $L; echo -n "N=$n no defer: " >> $L; MCB_SLAB_NO_DEFER=1 timeout 120 tools/dbg/dense_time $n 1 2>&1 | grep "build + inverse" >> $L; done
echo "== slab_prof" >> $L; timeout 120 tools/dbg/slab_prof >> $L 2>&1
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-steady --skip-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['roofline']['phase_ms_per_step']; print('C2', round(d['ms_per_step'],3), 'theta_i', round(p['mstep_indiv'],3), 'theta_k', round(p['mstep_mean'],3), 'us/eval', round(d['roofline']['all']['mean_process_evaluations']['us_per_evaluation'],1), 'frac', round(d['roofline']['frac'],4))
" >> $L
done
