#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_lb.log; : > $L
timeout 120 tools/dbg/eval_prof 50 40 >> $L 2>&1
timeout 120 tools/dbg/eval_prof 30 40 >> $L 2>&1
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-steady --skip-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['roofline']['phase_ms_per_step']; print('C2', round(d['ms_per_step'],3), 'theta_i', round(p['mstep_indiv'],3), 'theta_k', round(p['mstep_mean'],3), 'us/eval', round(d['roofline']['all']['mean_process_evaluations']['us_per_evaluation'],1), 'frac', round(d['roofline']['frac'],4))
" >> $L
done
MCB_NO_OVERLAP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-steady --skip-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['roofline']['phase_ms_per_step']; print('C2 no overlap', round(d['ms_per_step'],3), 'theta_i', round(p['mstep_indiv'],3), 'theta_k', round(p['mstep_mean'],3))
" >> $L
timeout 900 python -m pytest tests -m gpu -x -q -k "mstep or training or baseline" > gpurun_out/r02s3_lb_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02s3_lb_tests.log
