"""CPU: the EM driver of the free-hp branch of train_new_gp_EM (MC.R:142-186, repaired; csrc/new_indiv_em.cu) with the
ORACLE's log-density behind its callback. The driver is host code: everything but the GPU kernel that normally sits behind
the callback is exercised here (lock-step state machine over several new individuals, central differences, L-BFGS,
stopping rule), and compared with the oracle's own EM loop (scipy L-BFGS-B: trajectories are not comparable, end points are)."""
import ctypes as C

import numpy as np
import pytest

from oracle import magma_oracle as O
from helpers import small_problem
from magmaclust_b200 import _lib


def _setting(seed=9, n_new=(5, 8, 3)):
    db, hp, tau, m_k = small_problem(M=10, seed=seed, common_hp_i=True)
    param = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    rng = np.random.default_rng(seed)
    grid = np.unique(np.asarray(db["Timestamp"], dtype=np.float64))
    news = []
    for j, n in enumerate(n_new):
        t = np.sort(rng.choice(grid, size=n, replace=False))
        news.append({"ID": np.array([900 + j] * n), "Timestamp": t,
                     "Output": np.sin(t / 2.0 + j) + 0.1 * rng.standard_normal(n)})
    mu = O.posterior_mu_k(db, grid, m_k, O.kernel_mu, O.kernel, {"hp": hp, "param": param})
    return mu, news


def _run_driver(mu, news, theta0, n_loop_max=25, per_indiv=False):
    lib = _lib.load()
    names_k = list(mu["mean"].keys())
    K, B = len(names_k), len(news)
    pi_k = np.array([np.mean(list(mu["tau_i_k"][k].values())) for k in names_k])
    calls = {"n": 0, "pts": 0}

    def cb(ctx, npts, pt_ind, pt_theta, out):
        calls["n"] += 1
        calls["pts"] += npts
        for p in range(npts):
            th = np.array([pt_theta[3 * p + j] for j in range(3)])
            ll = O.new_indiv_logl(news[pt_ind[p]], mu["mean"], mu["cov"], O.kernel, th)
            for k in range(K):
                out[p * K + k] = ll[k]
        return 0

    fn = _lib.LOGL_FN(cb)
    theta0 = np.ascontiguousarray(theta0, dtype=np.float64)
    th, tau, loops = np.empty((B, 3)), np.empty((B, K)), np.empty(B, dtype=np.int32)
    rc = lib.mcb_new_indiv_em(B, K, _lib.dptr(theta0), int(per_indiv), _lib.dptr(pi_k), n_loop_max, fn, None,
                              _lib.dptr(th), _lib.dptr(tau), _lib.iptr(loops))
    assert rc == 0
    return th, tau, loops, pi_k, calls


def test_driver_matches_the_oracle_em():
    mu, news = _setting()
    theta0 = np.array([1.0, 1.0, 0.2])
    th, tau, loops, pi_k, calls = _run_driver(mu, news, theta0)
    assert calls["pts"] % 7 == 0 and calls["n"] >= loops.max()
    for b, new in enumerate(news):
        ref = O.train_new_gp_EM(new, mu, {"theta_i": theta0}, O.kernel, hp_i=None, return_loops=True)
        tau_ref = np.array([ref["tau_k"][k] for k in mu["mean"]])
        assert 1 <= loops[b] <= 25 and abs(tau[b].sum() - 1.0) < 1e-12
        assert tau[b].argmax() == tau_ref.argmax() and np.abs(tau[b] - tau_ref).max() < 5e-3
        # same optimum of the same objective: compare the objective values at the two end points with the same weights
        f_got = O.new_indiv_objective(th[b], tau_ref, new, mu["mean"], mu["cov"], O.kernel)
        f_ref = O.new_indiv_objective(ref["theta_new"], tau_ref, new, mu["mean"], mu["cov"], O.kernel)
        f_0 = O.new_indiv_objective(theta0, tau_ref, new, mu["mean"], mu["cov"], O.kernel)
        assert f_got <= f_0 and abs(f_got - f_ref) <= 1e-3 * max(1.0, abs(f_ref))


def test_returned_tau_is_the_e_step_of_the_previous_theta_and_one_pass_is_one_m_step():
    """MC.R:192 returns tau_k of the last E step (at the theta BEFORE the last M step). With n_loop_max = 1 that is the
    E step at theta0, and theta_new is one M step from it: the objective does not go up"""
    mu, news = _setting(seed=4, n_new=(6, 4))
    theta0 = np.array([[0.5, 0.8, 0.3], [1.2, 0.4, 0.15]])
    th, tau, loops, pi_k, _ = _run_driver(mu, news, theta0, n_loop_max=1, per_indiv=True)
    assert list(loops) == [1, 1]
    for b, new in enumerate(news):
        e = O.update_tau_star_k_EM(new, mu["mean"], mu["cov"], O.kernel, theta0[b], pi_k)
        assert np.abs(tau[b] - np.array([e[k] for k in mu["mean"]])).max() < 1e-12
        f0 = O.new_indiv_objective(theta0[b], tau[b], new, mu["mean"], mu["cov"], O.kernel)
        f1 = O.new_indiv_objective(th[b], tau[b], new, mu["mean"], mu["cov"], O.kernel)
        assert f1 < f0


def test_not_positive_definite_start_and_bad_arguments():
    mu, news = _setting(seed=2, n_new=(4,))
    lib = _lib.load()
    K = len(mu["mean"])

    def cb(ctx, npts, pt_ind, pt_theta, out):
        for i in range(npts * K):
            out[i] = float("nan")
        return 0

    fn = _lib.LOGL_FN(cb)
    th, tau, loops = np.empty((1, 3)), np.empty((1, K)), np.empty(1, dtype=np.int32)
    theta0, pi = np.array([1.0, 1.0, 0.2]), np.full(K, 1.0 / K)
    assert lib.mcb_new_indiv_em(1, K, _lib.dptr(theta0), 0, _lib.dptr(pi), 25, fn, None, _lib.dptr(th), _lib.dptr(tau),
                                _lib.iptr(loops)) == 0
    assert loops[0] == -1 and np.all(np.isnan(tau)) and np.array_equal(th[0], theta0)
    # a failing callback ends the run with its code; bad arguments are rejected
    bad = _lib.LOGL_FN(lambda *a: 7)
    assert lib.mcb_new_indiv_em(1, K, _lib.dptr(theta0), 0, _lib.dptr(pi), 25, bad, None, _lib.dptr(th), _lib.dptr(tau),
                                _lib.iptr(loops)) == 7
    assert lib.mcb_new_indiv_em(0, K, _lib.dptr(theta0), 0, _lib.dptr(pi), 25, fn, None, _lib.dptr(th), _lib.dptr(tau),
                                _lib.iptr(loops)) < 0


def test_analytic_log_density_and_the_reference_stopping_quirk():
    """K = 1 and log-density -|theta - c|^2 / 2: tau = 1, the M step lands on c (central differences are exact for a
    quadratic). MC.R:177 stops on 0 < eps < 1e-3: an M step that does not move at all (eps == 0) does NOT stop the loop,
    so the run ends at n_loop_max — mirrored, not repaired."""
    lib = _lib.load()
    c = np.array([[0.3, -1.2, 0.7], [2.0, 0.5, 0.1]])

    def cb(ctx, npts, pt_ind, pt_theta, out):
        for p in range(npts):
            th = np.array([pt_theta[3 * p + j] for j in range(3)])
            out[p] = -0.5 * float(np.sum((th - c[pt_ind[p]]) ** 2))
        return 0

    fn = _lib.LOGL_FN(cb)
    th, tau, loops = np.empty((2, 3)), np.empty((2, 1)), np.empty(2, dtype=np.int32)
    theta0, pi = np.array([1.0, 1.0, 0.2]), np.array([1.0])
    assert lib.mcb_new_indiv_em(2, 1, _lib.dptr(theta0), 0, _lib.dptr(pi), 6, fn, None, _lib.dptr(th), _lib.dptr(tau),
                                _lib.iptr(loops)) == 0
    assert np.abs(th - c).max() < 1e-6 and np.array_equal(tau, np.ones((2, 1)))
    assert 2 <= loops.min() and loops.max() <= 6


def test_batch_composition_does_not_change_results():
    """the B loops run in lock step but are independent: splitting the batch (how new individuals shard over GPUs — no
    exchange step) gives bit-identical hyper-parameters, responsibilities and pass counts"""
    mu, news = _setting(seed=6, n_new=(4, 7, 5))
    theta0 = np.array([1.0, 1.0, 0.2])
    th, tau, loops, _, _ = _run_driver(mu, news, theta0)
    th_a, tau_a, loops_a, _, _ = _run_driver(mu, news[:2], theta0)
    th_b, tau_b, loops_b, _, _ = _run_driver(mu, news[2:], theta0)
    assert np.array_equal(th, np.vstack([th_a, th_b])) and np.array_equal(tau, np.vstack([tau_a, tau_b]))
    assert list(loops) == list(loops_a) + list(loops_b)


def test_oracle_log_density_against_scipy():
    """the oracle's building block against an independent implementation of the Gaussian log-density"""
    from scipy.stats import multivariate_normal
    mu, news = _setting(seed=3, n_new=(6, 2))
    for new in news:
        for th in ([1.0, 1.0, 0.2], [-0.4, 1.7, 0.05]):
            got = O.new_indiv_logl(new, mu["mean"], mu["cov"], O.kernel, np.array(th))
            t = np.asarray(new["Timestamp"])
            for c, k in enumerate(mu["mean"]):
                mk = mu["mean"][k]
                mean = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], t)]
                d = t[:, None] - t[None, :]
                cov = np.exp(th[0] - 0.5 * np.exp(-th[1]) * d * d) + th[2] ** 2 * np.eye(len(t)) + mu["cov"][k].sub(t)
                ref = multivariate_normal(mean, cov).logpdf(np.asarray(new["Output"]))
                assert abs(got[c] - ref) <= 1e-9 * max(1.0, abs(ref))
