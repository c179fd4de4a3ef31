"""Small runs of the round-2 kernels for compute-sanitizer (memcheck): individuals on the global-memory path, k-means,
new individuals with n* > 64, the large-grid inverse (fused panel kernel; the TMA update is forced with MCB_TMA_MIN_N=0)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from helpers import small_problem
from magmaclust_b200 import magma, init_io
from magmaclust_b200.engine import Engine

db, hp, tau, m_k = small_problem(M=4, K=2, n=(90, 90), n_grid=120, seed=3, ragged=False)
keep = np.ones(db["ID"].shape[0], bool)
keep[np.nonzero(db["ID"] == 2)[0][20:]] = False          # one small individual among the big ones
db = {c: v[keep] for c, v in db.items()}
s = magma.Session(0)
got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=s)
new = magma.m_step_VEM(db, hp, got, magma.kernel_mu, magma.kernel, m_k, True, False, session=s)
print("big individuals ok", [float(v[0]) for v in new["theta_i"].values()])
grid = np.round(np.linspace(0, 10, 120), 3)
mu = magma.posterior_mu_k(db, grid, m_k, magma.kernel_mu, magma.kernel, {"hp": hp, "param": got}, session=s)
t = np.sort(np.random.default_rng(0).choice(grid, 100, replace=False))
nd = {"Timestamp": t, "Output": np.sin(t)}
tk = magma.update_tau_star_k_EM(nd, mu["mean"], mu["cov"], magma.kernel, np.array([0.7, 0.4, 0.35]), {k: 0.5 for k in m_k}, session=s)
pr = magma.pred_gp_clust(nd, grid[::5], mu, magma.kernel, {"theta_new": np.array([0.7, 0.4, 0.35]), "tau_k": tk}, session=s)
print("big new individual ok", tk)
rng = np.random.default_rng(1)
feat = rng.normal(0, 1, (1500, 3)) + 3.0 * rng.integers(0, 3, 1500)[:, None]
lab, cent, cost, it = s.eng.kmeans(feat, 3, init_io.kmeans_starts(1500, 3, 4, 0))
print("kmeans ok", cost, it)
x = np.linspace(0, 10, 800)
Kinv, ld = s.eng.kern_to_inv(x, np.array([1.0, 1.0]), 0.3)
print("large-grid inverse ok", ld, float(np.abs(Kinv - Kinv.T).max()))
s.close()
