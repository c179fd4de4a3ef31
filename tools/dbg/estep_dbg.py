import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from helpers import *
from magmaclust_b200 import magma
from oracle import magma_oracle as O
s = magma.Session(0)
db, hp, tau, m_k = gold_dataset(load_gold(), 0)
c = s.ensure_data(db)
print('M,N', c.M, c.N, 'nmax', np.diff(c.offsets).max())
names = list(m_k)
tk, ti = s.pack_hp(hp, names)
print(tk, ti[:2])
try:
    s.eng.set_state(tk, ti, hp['pi_k'], s.pack_mk(m_k, names), s.pack_tau(tau, names))
    s.eng.e_step()
    print('ok')
except Exception as e:
    print('ERR', e)
db, hp, tau, m_k = small_problem(seed=0)
try:
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=s)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau, return_logL=True)
    for k in m_k:
        print(k, rel_err(got['mean'][k]['Output'], ref['mean'][k]['Output']), rel_err(got['cov'][k].m, ref['cov'][k].m))
    print('tau', np.abs(O.tau_to_array(got['tau_i_k']) - O.tau_to_array(ref['tau_i_k'])).max())
except Exception as e:
    print('ERR small', e)
