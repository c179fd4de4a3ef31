#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --reps 1 --no-steady --no-extras --skip-cpu-baseline > gpurun_out/r02s2_launches_plain.json 2> gpurun_out/r02s2_launches_plain.err
echo "rc=$?" >> gpurun_out/r02s2_launches_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s2_launches.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-steady --no-extras --skip-cpu-baseline > gpurun_out/r02s2_launches_ncu.log 2>&1
echo done
