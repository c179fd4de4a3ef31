"""GPU edge cases: tiny and large individuals (all template instantiations of the evaluation kernel), K = 1, grids
smaller than one pivot block, error paths, batched prediction, full_algo_clust end to end."""
import numpy as np
import pytest

from oracle import magma_oracle as O
from helpers import RTOL_LL, RTOL_POST, assert_param_close, rel_err, small_problem

pytestmark = pytest.mark.gpu


def _check_eval(session, db, hp, tau, m_k, seed=0):
    from magmaclust_b200 import magma
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert_param_close(got, ref, list(m_k.keys()))
    per = O.split_db(db)
    rng = np.random.default_rng(seed)
    th = np.array([np.asarray(hp["theta_i"][i]) + 0.1 * rng.standard_normal(3) for i in per])
    f, g = session.eng.eval_indiv(th)
    for r, (i, d) in enumerate(per.items()):
        fr = O.logL_clust_multi_GP(th[r], d, ref, O.kernel)
        gr = O.gr_clust_multi_GP(th[r], d, ref, O.kernel)
        assert abs(f[r] - fr) <= RTOL_LL * max(1.0, abs(fr)), (i, f[r], fr)
        assert np.max(np.abs(g[r] - gr)) <= RTOL_LL * max(1.0, np.max(np.abs(gr))), (i, g[r], gr)
    return ref, got


@pytest.mark.parametrize("nrange,n_grid", [((1, 3), 12), ((20, 30), 60), ((31, 40), 90), ((38, 44), 90), ((45, 54), 100),
                                           ((55, 56), 100), ((60, 75), 120), ((78, 80), 120)])
def test_individual_sizes_cover_all_kernel_shapes(session, nrange, n_grid):
    """n_i from 1 (scalar branch, CF.R:167-170) to 80 (the shared-memory limit): every (threads, tile bucket, design)
    instantiation of the evaluation kernels (indiv.cu: eval2_shape)"""
    db, hp, tau, m_k = small_problem(M=6, K=2, n=nrange, n_grid=n_grid, seed=11)
    _check_eval(session, db, hp, tau, m_k)


@pytest.mark.parametrize("var", ["MCB_EVAL_V3", "MCB_EVAL_V2", "MCB_EVAL_V1"])
def test_other_evaluator_designs_agree_with_the_oracle(var):
    """the design is chosen once per process (environment variable read at the first launch): re-run the size sweep in a
    child process with the blocked-DMMA (V3), rank-1 (V2) and first (V1) evaluators forced for every bucket they support"""
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    env[var] = "1"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_edges.py"), "-q", "-x", "-m", "gpu",
                        "-k", "test_individual_sizes_cover_all_kernel_shapes"], env=env, cwd=root, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def _mixed_problem(M, K, sizes, n_grid, seed, common_hp_i=False):
    """a small problem whose individual r has sizes[r % len(sizes)] observations"""
    db, hp, tau, m_k = small_problem(M=M, K=K, n=(max(sizes), max(sizes)), n_grid=n_grid, seed=seed, ragged=False,
                                     common_hp_i=common_hp_i)
    keep = np.zeros(db["ID"].shape[0], dtype=bool)
    rng = np.random.default_rng(seed + 1)
    for r, i in enumerate(np.unique(db["ID"])):
        rows = np.nonzero(db["ID"] == i)[0]
        keep[np.sort(rng.choice(rows, sizes[r % len(sizes)], replace=False))] = True
    return {c: v[keep] for c, v in db.items()}, hp, tau, m_k


@pytest.mark.parametrize("sizes", [(81, 12, 130, 40), (96, 100), (200, 7)])
def test_more_than_80_observations_take_the_global_memory_kernels(session, sizes):
    """the reference has no limit on n_i (CF.R:153-190); above the tile buckets of the shared-memory kernels an individual
    runs on indiv_big.cu. Mixed with small individuals, all big, and n_i = 200: E-step, objective and gradient vs the oracle"""
    db, hp, tau, m_k = _mixed_problem(M=6, K=2, sizes=sizes, n_grid=260, seed=17)
    _check_eval(session, db, hp, tau, m_k)


@pytest.mark.parametrize("common_hp_i", [True, False])
def test_m_step_with_big_individuals_reaches_the_oracles_optimum(session, common_hp_i):
    """m_step_VEM (CF.R:631-687) on a data set with individuals on both paths: the optimiser (in-kernel for the small ones,
    the same state machine in lock step on the host for the big ones) ends at scipy's minimiser of the oracle's objective"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = _mixed_problem(M=5, K=2, sizes=(90, 14, 110), n_grid=160, seed=23, common_hp_i=common_hp_i)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    new = magma.m_step_VEM(db, hp, got, magma.kernel_mu, magma.kernel, m_k, True, common_hp_i, session=session)
    want = O.m_step_VEM(db, hp, ref, O.kernel_mu, O.kernel, m_k, True, common_hp_i)
    per = O.split_db(db)
    for i in per:
        a, b = np.asarray(new["theta_i"][i]), np.asarray(want["theta_i"][i])
        if common_hp_i:
            fa = O.logL_clust_multi_GP_common_hp_i(a, db, ref, O.kernel)
            fb = O.logL_clust_multi_GP_common_hp_i(b, db, ref, O.kernel)
        else:
            fa, fb = O.logL_clust_multi_GP(a, per[i], ref, O.kernel), O.logL_clust_multi_GP(b, per[i], ref, O.kernel)
        # same objective value at the end point (flat directions make theta itself loose), never worse than scipy's
        assert fa <= fb + 1e-6 * max(1.0, abs(fb)), (i, fa, fb)
        assert abs(fa - fb) <= 1e-4 * max(1.0, abs(fb)), (i, fa, fb)
    elbo_g = magma.logL_monitoring_VEM(new, db, magma.kernel, magma.kernel_mu, got, m_k, session=session)
    elbo_o = O.logL_monitoring_VEM(new, db, O.kernel, O.kernel_mu, ref, m_k)
    assert abs(elbo_g - elbo_o) <= RTOL_LL * abs(elbo_o)


def test_new_individual_with_more_than_64_observations(session):
    """pred_gp_clust / update_tau_star_k_EM / the free-hp log-density for n* = 100 and n* = 180 (MC.R:255-262 has no limit):
    the global-memory versions of the prediction kernels against the oracle"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=10, K=2, n=(8, 14), n_grid=200, seed=29, common_hp_i=True)
    names = list(m_k.keys())
    param = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    list_hp = {"hp": hp, "param": param}
    grid = np.round(np.linspace(0, 10, 200), 3)
    mu = magma.posterior_mu_k(db, grid, m_k, magma.kernel_mu, magma.kernel, list_hp, session=session)
    mu_o = O.posterior_mu_k(db, grid, m_k, O.kernel_mu, O.kernel, {"hp": hp, "param": ref})
    rng = np.random.default_rng(5)
    for nstar in (100, 180):
        t = np.sort(rng.choice(grid, nstar, replace=False))
        new_db = {"Timestamp": t, "Output": 1.0 + np.sin(t) + 0.3 * rng.standard_normal(nstar)}
        th = np.array([0.7, 0.4, 0.35])
        pi_k = {k: 1.0 / len(names) for k in names}
        tk = magma.update_tau_star_k_EM(new_db, mu["mean"], mu["cov"], magma.kernel, th, pi_k, session=session)
        tk_o = O.update_tau_star_k_EM(new_db, mu_o["mean"], mu_o["cov"], O.kernel, th, pi_k)
        for k in names:
            assert abs(tk[k] - tk_o[k]) < RTOL_LL
        pr = magma.pred_gp_clust(new_db, grid[::4], mu, magma.kernel, {"theta_new": th, "tau_k": tk}, session=session)
        pr_o = O.pred_gp_clust(new_db, grid[::4], mu_o, O.kernel, {"theta_new": th, "tau_k": tk_o})
        for k in names:
            assert rel_err(pr[k]["Mean"], pr_o[k]["Mean"]) < RTOL_POST
            assert np.max(np.abs(np.asarray(pr[k]["Var"]) - np.asarray(pr_o[k]["Var"]))) < RTOL_POST * max(1.0, np.max(np.abs(pr_o[k]["Var"])))
        idx = np.searchsorted(grid, t).astype(np.int32)
        ll = session.eng.new_indiv_logl(np.array([0, nstar], np.int32), idx, new_db["Output"], np.array([0], np.int32), th[None])
        ll_o = O.new_indiv_logl(new_db, mu_o["mean"], mu_o["cov"], O.kernel, th)
        assert np.max(np.abs(ll[0] - ll_o)) <= RTOL_LL * max(1.0, np.max(np.abs(ll_o)))


def test_single_cluster_and_tiny_grid(session):
    db, hp, tau, m_k = small_problem(M=5, K=1, n=(3, 6), n_grid=9, seed=2)
    ref, got = _check_eval(session, db, hp, tau, m_k)
    A = O.tau_to_array(got["tau_i_k"])
    assert np.allclose(A, 1.0)


def test_grid_sizes_around_pivot_block_boundaries(session):
    from magmaclust_b200 import magma
    for n_grid in (31, 32, 33, 64, 65):
        db, hp, tau, m_k = small_problem(M=14, K=2, n=(5, 9), n_grid=n_grid, seed=n_grid)
        ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
        got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
        assert_param_close(got, ref, list(m_k.keys()))


def test_not_positive_definite_is_reported_not_hidden(session):
    """the reference silently falls back to MASS::ginv (CF.R:142); the library reports MCB_ENOTPD"""
    from magmaclust_b200._lib import McbError
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=8, K=2, n=(6, 9), n_grid=60, seed=3)
    hp["theta_k"] = {k: np.array([9.0, 6.0, 0.0]) for k in hp["theta_k"]}  # smooth kernel, zero jitter: singular
    with pytest.raises(McbError) as ei:
        magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert ei.value.code == 2
    # the handle stays usable afterwards
    db, hp, tau, m_k = small_problem(M=8, K=2, n=(6, 9), n_grid=60, seed=3)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert_param_close(got, ref, list(m_k.keys()))


def test_negative_noise_sd_is_not_clamped(session):
    """theta[3] may be negative: only its square enters Psi, the gradient keeps the sign"""
    db, hp, tau, m_k = small_problem(M=6, K=2, n=(5, 8), n_grid=30, seed=4)
    hp["theta_i"] = {i: np.array([v[0], v[1], -abs(v[2])]) for i, v in hp["theta_i"].items()}
    _check_eval(session, db, hp, tau, m_k)


def test_batched_prediction_matches_per_individual_oracle(session):
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=10, K=3, seed=12, common_hp_i=True)
    names = list(m_k.keys())
    ref_param = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    list_hp = {"hp": hp, "param": ref_param}
    grid = np.unique(np.r_[db["Timestamp"], np.round(np.linspace(0.05, 9.95, 37), 3)])
    mu_ref = O.posterior_mu_k(db, grid, m_k, O.kernel_mu, O.kernel, list_hp)
    magma.posterior_mu_k(db, grid, m_k, magma.kernel_mu, magma.kernel, list_hp, session=session)
    rng = np.random.default_rng(0)
    B = 7
    offs, idx, ys, news = [0], [], [], []
    for b in range(B):
        nb = int(rng.integers(1, 9))
        sel = np.sort(rng.choice(grid.shape[0], nb, replace=False))
        y = rng.standard_normal(nb)
        idx += list(sel)
        ys += list(y)
        offs.append(offs[-1] + nb)
        news.append({"ID": np.array([100 + b] * nb), "Timestamp": grid[sel], "Output": y})
    theta_new = np.asarray(next(iter(hp["theta_i"].values())))
    pred_idx = np.arange(0, grid.shape[0], 3)
    pi_k = np.array([np.mean(list(tau[k].values())) for k in names])
    tau_b, mean_b, var_b = session.eng.predict(offs, idx, ys, theta_new, pred_idx, pi_k)
    for b in range(B):
        t_ref = O.update_tau_star_k_EM(news[b], mu_ref["mean"], mu_ref["cov"], O.kernel, theta_new,
                                       {k: pi_k[c] for c, k in enumerate(names)})
        p_ref = O.pred_gp_clust(news[b], grid[pred_idx], mu_ref, O.kernel, {"theta_new": theta_new, "tau_k": t_ref})
        for c, k in enumerate(names):
            assert abs(tau_b[b, c] - t_ref[k]) < RTOL_LL
            assert rel_err(mean_b[b, c], p_ref[k]["Mean"]) < RTOL_POST
            assert np.max(np.abs(var_b[b, c] - p_ref[k]["Var"])) < 1e-8 * max(1.0, np.max(np.abs(p_ref[k]["Var"])))


def test_full_algo_clust_end_to_end(session):
    """MC.R:301-347 with list_hp = NULL: training, hyper-posterior on the prediction grid, tau*, prediction"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=12, K=2, n=(8, 12), n_grid=40, seed=21, common_hp_i=True)
    ids = list(O.split_db(db).keys())
    names = list(m_k.keys())
    A = O.tau_to_array(tau)
    tau0 = {k: {i: 1.0 if A[r].argmax() == c else 0.0 for r, i in enumerate(ids)} for c, k in enumerate(names)}
    ini = {"theta_k": np.array([1.0, 1.0, 0.2]), "theta_i": np.array([1.0, 1.0, 0.2])}
    new_db = {"ID": np.array([99] * 4), "Timestamp": np.array([0.513, 3.077, 5.128, 8.205]), "Output": np.array([0.1, 0.8, -0.3, 0.5])}
    targets = np.round(np.linspace(0.2, 9.8, 11), 3)
    ref = O.full_algo_clust(db, new_db, targets, O.kernel, tau0, True, True, prior_mean=m_k, kern_0=O.kernel_mu, ini_hp=ini)
    got = magma.full_algo_clust(db, new_db, targets, magma.kernel, tau0, True, True, prior_mean=m_k,
                                kern_0=magma.kernel_mu, ini_hp=ini, session=session)
    assert len(got["logL_iterations"]) == len(ref["logL_iterations"])
    for k in names:
        assert abs(got["Hyperparameters"]["new_i"]["tau_k"][k] - ref["Hyperparameters"]["new_i"]["tau_k"][k]) < 1e-3
        assert np.allclose(got["Prediction"][k]["Mean"], ref["Prediction"][k]["Mean"], rtol=2e-3, atol=2e-3)
        assert np.allclose(got["Prediction"][k]["Var"], ref["Prediction"][k]["Var"], rtol=5e-3, atol=1e-4)


@pytest.mark.parametrize("n", [1, 2, 7, 33, 120, 801])
def test_standalone_kern_to_cov_inv_dmvnorm(session, n):
    """the vector branches of kern_to_cov / kern_to_inv (CF.R:61-149) and dmvnorm (CF.R:10-35) as their own entry points
    (n = 801 takes the multi-launch inverse)"""
    from magmaclust_b200 import magma
    rng = np.random.default_rng(n)
    x = rng.permutation(np.round(np.sort(rng.uniform(0, 10, n)), 6))        # unsorted on purpose: both sides sort
    theta, sigma = (0.8, 0.3), 0.45
    c_ref, c_got = O.kern_to_cov(x, O.kernel, theta, sigma), magma.kern_to_cov(x, magma.kernel, theta, sigma, session=session)
    assert np.array_equal(c_ref.t, c_got.t)
    assert rel_err(c_got.m, c_ref.m) < 1e-14
    i_ref, i_got = O.kern_to_inv(x, O.kernel, theta, sigma), magma.kern_to_inv(x, magma.kernel, theta, sigma, session=session)
    assert rel_err(i_got.m, i_ref.m) < RTOL_POST
    mu = rng.normal(size=n)
    xv = mu + 0.3 * rng.normal(size=n)
    for lg in (True, False):
        d_ref = O.dmvnorm(xv, mu, i_ref.m, log=lg)
        d_got = magma.dmvnorm(xv, mu, i_got, log=lg, session=session)
        assert abs(d_got - d_ref) <= RTOL_LL * max(abs(d_ref), 1e-300)
    assert np.array_equal(magma.mat_dist(x[:5], x[:3]), O.mat_dist(x[:5], x[:3]))
    assert np.array_equal(magma.deriv_hp2(O.mat_dist(x, x), theta), O.deriv_hp2(O.mat_dist(x, x), theta))


def test_standalone_kern_to_inv_db_branch(session):
    """db branch: one matrix per ID with theta[[i]][3] as the noise sd (CF.R:153-190)"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=4, K=2, n=(4, 9), n_grid=30, seed=9)
    ref = O.kern_to_inv(db, O.kernel, hp["theta_i"])
    got = magma.kern_to_inv(db, magma.kernel, hp["theta_i"], session=session)
    assert list(ref.keys()) == list(got.keys())
    for i in ref:
        assert rel_err(got[i].m, ref[i].m) < RTOL_POST
