#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s3_tests_exp.log 2>&1
echo "rc=$?" >> gpurun_out/r02s3_tests_exp.log
timeout 900 python bench.py > gpurun_out/r02s3_bench_exp.json 2> gpurun_out/r02s3_bench_exp.err
echo "rc=$?" >> gpurun_out/r02s3_bench_exp.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s3_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r02s3_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02s3_bench_ref.json 2> gpurun_out/r02s3_bench_ref.err
echo "rc=$?" >> gpurun_out/r02s3_bench_ref.err
