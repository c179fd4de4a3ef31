#!/bin/bash
# round 2, GPU call 2: ncu --set full with source of the per-individual M-step kernel (theta_i alone) and of the slab inverse
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out

MCB_NO_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"indiv_mstep2_kernel" -c 1 -s 2 \
  -o gpurun_out/r02_mstep2_alone python bench.py --steps 2 --warmup 1 --reps 1 --no-steady --skip-cpu-baseline > gpurun_out/r02_exp2_ncu.log 2>&1
MCB_HOST_THETA_K=1 MCB_NO_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spd_inverse_slab_kernel" -c 1 -s 12 \
  -o gpurun_out/r02_slab python bench.py --steps 2 --warmup 1 --reps 1 --no-steady --skip-cpu-baseline > gpurun_out/r02_exp2_ncu2.log 2>&1
echo done
