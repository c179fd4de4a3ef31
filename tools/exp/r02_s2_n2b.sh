#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s2_tests7.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_tests7.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02s2_bench_n2b.json 2> gpurun_out/r02s2_bench_n2b.err
echo "rc=$?" >> gpurun_out/r02s2_bench_n2b.err
