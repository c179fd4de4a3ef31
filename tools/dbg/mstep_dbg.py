import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from helpers import *
from magmaclust_b200 import magma
from oracle import magma_oracle as O
s = magma.Session(0)
db, hp, tau, m_k = small_problem(M=10, seed=4, common_hp_i=False)
first = next(iter(hp["theta_k"].values()))
hp["theta_k"] = {k: first.copy() for k in hp["theta_k"]}
got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=s)
tk0, ti0, _ = s.eng.get_hp()
f0, g0 = s.eng.eval_indiv(ti0)
tk, ti, pi, st = s.eng.m_step(True, False)
print(st)
f1, g1 = s.eng.eval_indiv(ti)
for r in range(len(ti0)):
    print(r, ti0[r], '->', ti[r], 'f', f0[r], '->', f1[r], 'g1', g1[r])
print('theta_k', tk0[0], '->', tk[0])
ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
per = O.split_db(db)
from scipy.optimize import minimize
for r, (i, d) in enumerate(per.items()):
    res = minimize(O.logL_clust_multi_GP, ti0[r], args=(d, ref, O.kernel), jac=O.gr_clust_multi_GP, method='L-BFGS-B', options={'maxcor':5,'ftol':2.2e-9,'gtol':0,'maxiter':100})
    print(r, res.x, res.fun, res.nit, res.nfev)
