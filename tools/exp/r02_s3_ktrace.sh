#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
MCB_TRACE_K=1 timeout 300 python tools/dbg/nfev.py > gpurun_out/r02s3_ktrace.log 2>&1
MCB_TRACE_K=1 MCB_NO_OVERLAP=1 timeout 300 python tools/dbg/nfev.py > gpurun_out/r02s3_ktrace_alone.log 2>&1
