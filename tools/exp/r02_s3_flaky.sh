#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r02s3_flaky.log; : > $L
cat > /tmp/k9.py <<'P'
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import small_problem
from magmaclust_b200 import magma
from oracle import magma_oracle as O
db, hp, tau, m_k = small_problem(M=27, K=9, seed=6, common_hp_i=False)
first = next(iter(hp["theta_k"].values()))
hp["theta_k"] = {k: first.copy() for k in hp["theta_k"]}
ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
new_ref = O.m_step_VEM(db, hp, ref, O.kernel_mu, O.kernel, m_k, True, False)
th_r = next(iter(new_ref["theta_k"].values()))
f_r = O.logL_GP_mod_common_hp_k(th_r[:2], ref["mean"], m_k, O.kernel_mu, ref["cov"], th_r[2])
for rep in range(12):
    s = magma.Session(0)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=s)
    new_got = magma.m_step_VEM(db, hp, got, magma.kernel_mu, magma.kernel, m_k, True, False, session=s)
    th_g = next(iter(new_got["theta_k"].values()))
    f_g = O.logL_GP_mod_common_hp_k(th_g[:2], ref["mean"], m_k, O.kernel_mu, ref["cov"], th_g[2])
    st = s.eng.last_mstep_stats()
    print(rep, "rel", (f_g - f_r) / abs(f_r), "theta_k", th_g, "ref", th_r, "nfev_k", st["nfev_k"], "nit_k", st["nit_k"], flush=True)
    s.close()
P
echo "== new lbfgs" >> $L; timeout 600 python /tmp/k9.py >> $L 2>&1
echo "== generic lbfgs" >> $L; MCB_LIB=$PWD/build/variants/libmcb_any.so timeout 600 python /tmp/k9.py >> $L 2>&1
