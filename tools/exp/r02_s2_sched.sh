#!/bin/bash
# scheduling sweep of the per-individual M-step kernel at C2 (priority SMs keeping 1 / 2 / 3 CTAs for the longest optimisations)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C tests/r_stub > /dev/null 2>&1
timeout 200 python tools/dbg/dump_nfev.py > gpurun_out/r02s2_nfev.json 2>&1
L=gpurun_out/r02s2_sched.log
: > $L
run() { echo "== $*" >> $L; env "$@" timeout 200 python tools/bench_phases.py --steps 20 --warmup 3 --no-steady >> $L 2>&1; }
run MCB_NOP=1
run MCB_NO_OVERLAP=1
for ct in 1 2 3; do for sm in 16 32 48 64; do for mult in 1 2; do
  run MCB_SOLO_CTAS=$ct MCB_SOLO_SMS=$sm MCB_SOLO_MULT=$mult
  run MCB_NO_OVERLAP=1 MCB_SOLO_CTAS=$ct MCB_SOLO_SMS=$sm MCB_SOLO_MULT=$mult
done; done; done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s2_tests3.log 2>&1
echo "rc=$?" >> gpurun_out/r02s2_tests3.log
