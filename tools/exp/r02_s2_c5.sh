#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_edges.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r02s2_tests4.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_tests4.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-steady > gpurun_out/r02s2_bench_full.json 2> gpurun_out/r02s2_bench_full.err
echo "rc=$?" >> gpurun_out/r02s2_bench_full.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"indiv_eval2_kernel" -c 1 -s 3 \
  -o gpurun_out/r02s2_eval2_c3 python tools/bench_extra.py c3 > gpurun_out/r02s2_ncu_eval2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"predict_kernel" -c 1 -s 1 \
  -o gpurun_out/r02s2_predict_c5 python tools/bench_extra.py c5 > gpurun_out/r02s2_ncu_predict.log 2>&1
echo done
