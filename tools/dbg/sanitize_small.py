"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every evaluator bucket, E-step, one VEM iteration."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import small_problem
from magmaclust_b200 import magma

for nrange, n_grid in (((5, 30), 60), ((45, 54), 100), ((60, 75), 120)):
    db, hp, tau, m_k = small_problem(M=6, K=2, n=nrange, n_grid=n_grid, seed=11)
    s = magma.Session(0)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=s)
    th = np.array([np.asarray(hp["theta_i"][i]) for i in hp["theta_i"]])
    f, g = s.eng.eval_indiv(th)
    out = s.eng.vem_iteration(True, False)
    print(nrange, float(f.sum()), out[:2])
    s.close()
