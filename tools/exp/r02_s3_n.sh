#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02s3_bench_n$N.json 2> gpurun_out/r02s3_bench_n$N.err
echo "rc=$?" >> gpurun_out/r02s3_bench_n$N.err
