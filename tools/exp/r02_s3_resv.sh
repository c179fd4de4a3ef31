#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_resv2.log; : > $L
for cfg in "26 0" "26 4" "26 8" "26 12" "32 4" "32 8" "22 8" "22 12" "39 0" "39 6" "26 8"; do
set -- $cfg
echo "== extra $1 solo $2" >> $L
MCB_RESERVE_EXTRA=$1 MCB_SOLO_SMS=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-steady --skip-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['roofline']['phase_ms_per_step']; print(round(d['ms_per_step'],3), 'theta_i', round(p['mstep_indiv'],3), 'theta_k', round(p['mstep_mean'],3), 'us/eval', round(d['roofline']['all']['mean_process_evaluations']['us_per_evaluation'],1))
" >> $L
done
