#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-1}
for v in "MCB_NOP=1" "MCB_HOST_THETA_I=1"; do
  tag=$(echo $v | cut -d= -f1)
  if [ "$N" = "1" ]; then
    env $v timeout 900 python bench.py --steps 6 --warmup 3 --no-steady --skip-cpu-baseline > gpurun_out/r02s2_iopt_n${N}_$tag.json 2> gpurun_out/r02s2_iopt_n${N}_$tag.err
  else
    env $v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02s2_iopt_n${N}_$tag.json 2> gpurun_out/r02s2_iopt_n${N}_$tag.err
  fi
  echo "rc=$?" >> gpurun_out/r02s2_iopt_n${N}_$tag.err
done
