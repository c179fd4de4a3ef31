"""GPU: new individuals with free hyper-parameters — the free-hp branch of train_new_gp_EM (MC.R:142-186, repaired, SURVEY
§8f row 3) through the C ABI, against the oracle's restatement. "Parity unpinned" for the EM end points (upstream cannot
run this branch and the optimisers differ); the log-densities themselves are compared at 1e-6 like every log-likelihood."""
import numpy as np
import pytest

from oracle import magma_oracle as O
from helpers import RTOL_LL, small_problem

pytestmark = pytest.mark.gpu


def _setting(session, seed=9, n_new=(5, 8, 3, 1)):
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=10, seed=seed, common_hp_i=True)
    param = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    list_hp = {"hp": hp, "param": param}
    rng = np.random.default_rng(seed)
    grid = np.unique(np.asarray(db["Timestamp"], dtype=np.float64))
    news = []
    for j, n in enumerate(n_new):
        t = np.sort(rng.choice(grid, size=n, replace=False))
        news.append({"ID": np.array([900 + j] * n), "Timestamp": t,
                     "Output": np.sin(t / 2.0 + j) + 0.1 * rng.standard_normal(n)})
    mu_ref = O.posterior_mu_k(db, grid, m_k, O.kernel_mu, O.kernel, list_hp)
    mu_got = magma.posterior_mu_k(db, grid, m_k, magma.kernel_mu, magma.kernel, list_hp, session=session)
    return mu_ref, mu_got, news, grid


def test_new_indiv_log_densities(session):
    from magmaclust_b200 import magma
    mu_ref, mu_got, news, grid = _setting(session)
    off, idx, y = magma._csr_new(news, grid)
    rng = np.random.default_rng(3)
    pt_ind = np.array([0, 1, 2, 3, 1, 0, 2, 3, 1], dtype=np.int32)
    pt_theta = np.column_stack([rng.uniform(-1.0, 2.0, 9), rng.uniform(-1.0, 2.0, 9), rng.uniform(0.05, 0.6, 9)])
    got = session.eng.new_indiv_logl(off, idx, y, pt_ind, pt_theta)
    for p in range(9):
        ref = O.new_indiv_logl(news[pt_ind[p]], mu_ref["mean"], mu_ref["cov"], O.kernel, pt_theta[p])
        assert np.abs(got[p] - ref).max() <= RTOL_LL * max(1.0, np.abs(ref).max())
    # the E step of mcb_predict is made of the same numbers
    th = pt_theta[4]
    pi = np.array([np.mean(list(mu_ref["tau_i_k"][k].values())) for k in mu_ref["mean"]])
    tau, _, _ = session.eng.predict(off, idx, y, th, None, pi)
    ll = session.eng.new_indiv_logl(off, idx, y, np.arange(4, dtype=np.int32), np.tile(th, (4, 1)))
    w = pi * np.exp(ll - ll.max(1, keepdims=True))
    assert np.abs(w / w.sum(1, keepdims=True) - tau).max() < 1e-12
    # a trial point that is not a covariance gives NaN for that point only, not an error
    bad = session.eng.new_indiv_logl(off, idx, y, np.array([0, 1], dtype=np.int32), np.array([[np.nan, 1.0, 0.2], [1.0, 1.0, 0.2]]))
    assert np.all(np.isnan(bad[0])) and np.all(np.isfinite(bad[1]))


def test_train_new_gp_EM_free_hyper_parameters(session):
    from magmaclust_b200 import magma
    mu_ref, mu_got, news, grid = _setting(session, n_new=(5, 8, 3))
    ini = {"theta_i": np.array([1.0, 1.0, 0.2])}
    got_all = magma.train_new_gp_EM(news, mu_got, ini, magma.kernel, hp_i=None, session=session)
    assert isinstance(got_all, list) and len(got_all) == 3
    for new, got in zip(news, got_all):
        ref = O.train_new_gp_EM(new, mu_ref, ini, O.kernel, hp_i=None, return_loops=True)
        tau_ref = np.array([ref["tau_k"][k] for k in mu_ref["mean"]])
        tau_got = np.array([got["tau_k"][k] for k in mu_ref["mean"]])
        assert 1 <= got["loops"] <= 25 and abs(tau_got.sum() - 1.0) < 1e-12
        assert tau_got.argmax() == tau_ref.argmax() and np.abs(tau_got - tau_ref).max() < 5e-3
        f_got = O.new_indiv_objective(got["theta_new"], tau_ref, new, mu_ref["mean"], mu_ref["cov"], O.kernel)
        f_ref = O.new_indiv_objective(ref["theta_new"], tau_ref, new, mu_ref["mean"], mu_ref["cov"], O.kernel)
        assert abs(f_got - f_ref) <= 1e-3 * max(1.0, abs(f_ref))
    # one new individual, not in a list: a single result, identical to its entry of the batch
    one = magma.train_new_gp_EM(news[1], mu_got, ini, magma.kernel, hp_i=None, session=session)
    assert np.array_equal(one["theta_new"], got_all[1]["theta_new"]) and one["tau_k"] == got_all[1]["tau_k"]
    # and the result feeds pred_gp_clust like the fixed-hp branch's
    pred = magma.pred_gp_clust(news[1], grid[::3], mu_got, magma.kernel, one, session=session)
    ref_pred = O.pred_gp_clust(news[1], grid[::3], mu_ref, O.kernel, {"theta_new": one["theta_new"], "tau_k": one["tau_k"]})
    for k in mu_ref["mean"]:
        assert np.abs(pred[k]["Mean"] - ref_pred[k]["Mean"]).max() <= 1e-8 * max(1.0, np.abs(ref_pred[k]["Mean"]).max())
