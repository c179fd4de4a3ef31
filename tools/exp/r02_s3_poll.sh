#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_poll.log; : > $L
timeout 120 tools/dbg/dense_check > gpurun_out/r02s3_dc.log 2>&1; rc=$?
echo "dense_check rc=$rc worst $(grep spd_inverse gpurun_out/r02s3_dc.log | awk '{print $6}' | sort -g | tail -1) info $(grep spd_inverse gpurun_out/r02s3_dc.log | awk '{print $NF}' | sort -u | tr '\n' ' ')" >> $L
if [ $rc -ne 0 ]; then exit 1; fi
for ns in 0 20 50 100 200 400; do for nd in 0 1; do
if [ $nd = 1 ]; then export MCB_SLAB_NO_DEFER=1; else unset MCB_SLAB_NO_DEFER; fi
echo -n "poll_ns=$ns no_defer=$nd: " >> $L; MCB_SLAB_POLL_NS=$ns timeout 120 tools/dbg/dense_time 401 1 2>&1 | grep "build + inverse" | tr '\n' ' ' >> $L; MCB_SLAB_POLL_NS=$ns timeout 120 tools/dbg/dense_time 401 5 2>&1 | grep "build + inverse" >> $L
done; done
unset MCB_SLAB_NO_DEFER
echo "== slab_prof poll 100" >> $L; MCB_SLAB_POLL_NS=100 timeout 120 tools/dbg/slab_prof >> $L 2>&1
