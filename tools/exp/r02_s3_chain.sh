#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_chain.log
echo "== dense_check" > $L
timeout 120 tools/dbg/dense_check > gpurun_out/r02s3_dc.log 2>&1; rc=$?; grep -v gemm_nt gpurun_out/r02s3_dc.log >> $L; echo "dense_check rc=$rc" >> $L
if [ $rc -ne 0 ]; then exit 1; fi
for z in 1 5; do echo "== dense_time 401 $z" >> $L; timeout 120 tools/dbg/dense_time 401 $z >> $L 2>&1; done
echo "== dense_time 768 5" >> $L; timeout 120 tools/dbg/dense_time 768 5 >> $L 2>&1
make -C tests/r_stub > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s3_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02s3_tests.log
timeout 900 python bench.py > gpurun_out/r02s3_bench.json 2> gpurun_out/r02s3_bench.err
echo "rc=$?" >> gpurun_out/r02s3_bench.err
