#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > /dev/null 2>&1
for i in 1 2; do
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02s3_repeat_$i.log 2>&1
echo "rc=$?" >> gpurun_out/r02s3_repeat_$i.log
done
