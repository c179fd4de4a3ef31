#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02s2_bench_n${N}c.json 2> gpurun_out/r02s2_bench_n${N}c.err
echo "rc=$?" >> gpurun_out/r02s2_bench_n${N}c.err
MCB_NO_PEER_ALLREDUCE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02s2_bench_n${N}c_nccl.json 2> gpurun_out/r02s2_bench_n${N}c_nccl.err
echo "rc=$?" >> gpurun_out/r02s2_bench_n${N}c_nccl.err
