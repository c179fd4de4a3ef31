#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
MCB_NO_OVERLAP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"indiv_mstep2_kernel" -c 1 -s 2 -f \
  -o gpurun_out/r02s3_mstep2_alone python bench.py --steps 2 --warmup 1 --reps 1 --no-steady --skip-cpu-baseline --no-extras > gpurun_out/r02s3_ncu_mstep2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02s3_launches.csv \
  python bench.py --steps 2 --warmup 1 --reps 1 --no-steady --skip-cpu-baseline --no-extras > gpurun_out/r02s3_ncu_launches.log 2>&1
echo done
