"""debug: is one VEM iteration reproducible between a fresh handle and a reused one (and between two fresh ones)?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from magmaclust_b200 import synth
from magmaclust_b200.engine import Engine

K = 3
a = synth.make_dataset(M=150, K=K, n=20, n_grid=81, common_hp_i=False, seed=15)
b = synth.make_dataset(M=150, K=K, n=20, n_grid=81, common_hp_i=False, seed=16)
tk = np.tile(np.array([1.0, 1.0, 0.2]), (K, 1)); ti = np.tile(np.array([1.0, 1.0, 0.2]), (150, 1)); mk = np.zeros(K)

def train(e, ds):
    e.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
    e.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
    r = e.vem_iteration(True, False)
    return r, e.get_new_hp(), e.last_mstep_stats()

res = []
for trial in range(3):
    e = Engine(0)
    res.append(("fresh%d" % trial, train(e, b)))
    e.close()
e = Engine(0)
train(e, a)
res.append(("after_a", train(e, b)))
res.append(("again", train(e, b)))
big = np.unique(np.r_[a["grid_full"], np.linspace(0.0, 10.0, 801)])
e.posterior_mu_k(big, tau=b["tau0"], want_mean=False, want_cov=False)
res.append(("after_post", train(e, b)))
e.close()
r0 = res[0][1]
for name, (r, hp, st) in res:
    dk = np.abs(hp[0] - r0[1][0]).max(); di = np.abs(hp[1] - r0[1][1]); 
    print(name, "elbo %.12g eps %.10g conv %s" % r, "theta_k", hp[0][0], "max dtheta_k %.3g max dtheta_i %.3g (indiv %d)" % (dk, di.max(), int(di.max(1).argmax())), "stats", st)
