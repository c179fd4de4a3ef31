#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_sens.log; : > $L
for nd in 0 1; do for dbg in 0 1 2 4 8 16; do
if [ $nd = 1 ]; then export MCB_SLAB_NO_DEFER=1; else unset MCB_SLAB_NO_DEFER; fi
echo -n "no_defer=$nd dbg=$dbg: " >> $L; MCB_SLAB_DBG=$dbg timeout 120 tools/dbg/dense_time 401 1 2>&1 | grep "build + inverse" >> $L
done; done
for nd in 0 1; do for ns in 0 50 200; do
if [ $nd = 1 ]; then export MCB_SLAB_NO_DEFER=1; else unset MCB_SLAB_NO_DEFER; fi
echo -n "no_defer=$nd poll_ns=$ns: " >> $L; MCB_SLAB_POLL_NS=$ns timeout 120 tools/dbg/dense_time 401 1 2>&1 | grep "build + inverse" | tr '\n' ' ' >> $L; MCB_SLAB_POLL_NS=$ns timeout 120 tools/dbg/dense_time 401 5 2>&1 | grep "build + inverse" >> $L
done; done
for q in 2 3 4; do for nd in 0 1; do
if [ $nd = 1 ]; then export MCB_SLAB_NO_DEFER=1; else unset MCB_SLAB_NO_DEFER; fi
echo -n "Q=$q no_defer=$nd: " >> $L; MCB_SLAB_Q=$q timeout 120 tools/dbg/dense_time 401 1 2>&1 | grep "build + inverse" >> $L
done; done
