"""GPU parity of the M-step callbacks (CF.R:194-325, 448-586), the ELBO monitor (CF.R:327-352, MC.R:51-53)
and the built-in optimiser (m_step_VEM, CF.R:631-687) against the CPU oracle."""
import numpy as np
import pytest

from oracle import magma_oracle as O
from helpers import NAMES3, RTOL_LL, gold_dataset, load_gold, rel_err, small_problem

pytestmark = pytest.mark.gpu


def _setup(session, db, hp, tau, m_k):
    from magmaclust_b200 import magma
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    return ref, got


@pytest.mark.parametrize("case", ["gold0", "small0", "small3"])
def test_indiv_objective_and_gradient(session, case):
    from magmaclust_b200 import magma
    if case.startswith("gold"):
        db, hp, tau, m_k = gold_dataset(load_gold(), 0)
    else:
        db, hp, tau, m_k = small_problem(seed=int(case[-1]))
    ref, got = _setup(session, db, hp, tau, m_k)
    per_id = O.split_db(db)
    rng = np.random.default_rng(3)
    th = np.array([np.asarray(hp["theta_i"][i]) + 0.2 * rng.standard_normal(3) for i in per_id])
    f, g = session.eng.eval_indiv(th)
    for r, (i, d) in enumerate(per_id.items()):
        fr = O.logL_clust_multi_GP(th[r], d, ref, O.kernel)
        gr = O.gr_clust_multi_GP(th[r], d, ref, O.kernel)
        assert abs(f[r] - fr) <= RTOL_LL * max(1.0, abs(fr))
        assert np.max(np.abs(g[r] - gr)) <= RTOL_LL * max(1.0, np.max(np.abs(gr)))
    # shared theta: CF.R:280-325 / 539-586
    th0 = np.array([0.9, 0.6, 0.4])
    fc, gc = session.eng.eval_indiv_common(th0)
    frc = O.logL_clust_multi_GP_common_hp_i(th0, db, ref, O.kernel)
    grc = O.gr_clust_multi_GP_common_hp_i(th0, db, ref, O.kernel)
    assert abs(fc - frc) <= RTOL_LL * abs(frc)
    assert np.max(np.abs(gc - grc)) <= RTOL_LL * max(1.0, np.max(np.abs(grc)))
    # R-shaped single-individual callbacks
    i0 = next(iter(per_id))
    assert abs(magma.logL_clust_multi_GP(th[0], per_id[i0], got, magma.kernel, session=session) -
               O.logL_clust_multi_GP(th[0], per_id[i0], ref, O.kernel)) <= RTOL_LL * abs(f[0])


@pytest.mark.parametrize("case", ["gold1", "small1"])
def test_mean_objective_and_gradient(session, case):
    if case.startswith("gold"):
        db, hp, tau, m_k = gold_dataset(load_gold(), 1)
    else:
        db, hp, tau, m_k = small_problem(seed=1)
    ref, got = _setup(session, db, hp, tau, m_k)
    names = list(m_k.keys())
    th3 = np.array([1.4, 0.9, 0.35])
    for ck, k in enumerate(names):
        for nhp in (2, 3):
            f, g = session.eng.eval_mean(ck, th3[:nhp], pen_diag=0.27)
            fr = O.logL_GP_mod(th3[:nhp], ref["mean"][k], m_k[k], O.kernel_mu, ref["cov"][k], 0.27)
            gr = O.gr_GP_mod(th3[:nhp], ref["mean"][k], m_k[k], O.kernel_mu, ref["cov"][k], 0.27)
            assert abs(f - fr) <= RTOL_LL * max(1.0, abs(fr))
            assert np.max(np.abs(g - gr)) <= RTOL_LL * max(1.0, np.max(np.abs(gr)))
    for nhp in (2, 3):
        f, g = session.eng.eval_mean_common(th3[:nhp], pen_diag=0.27)
        fr = O.logL_GP_mod_common_hp_k(th3[:nhp], ref["mean"], m_k, O.kernel_mu, ref["cov"], 0.27)
        gr = O.gr_GP_mod_common_hp_k(th3[:nhp], ref["mean"], m_k, O.kernel_mu, ref["cov"], 0.27)
        assert abs(f - fr) <= RTOL_LL * max(1.0, abs(fr))
        assert np.max(np.abs(g - gr)) <= RTOL_LL * max(1.0, np.max(np.abs(gr)))


@pytest.mark.parametrize("case", ["gold0", "small2"])
def test_elbo(session, case):
    from magmaclust_b200 import magma
    if case.startswith("gold"):
        db, hp, tau, m_k = gold_dataset(load_gold(), 0)
    else:
        db, hp, tau, m_k = small_problem(seed=2)
    ref, got = _setup(session, db, hp, tau, m_k)
    l0, l1 = session.eng.elbo()
    r0 = O.logL_monitoring_VEM(hp, db, O.kernel, O.kernel_mu, ref, m_k)
    r1 = O.elbo_VEM(hp, db, O.kernel, O.kernel_mu, ref, m_k)
    assert abs(l0 - r0) <= RTOL_LL * abs(r0)
    assert abs(l1 - r1) <= RTOL_LL * abs(r1)
    # at another hp, through the R-shaped call
    hp2 = {"theta_k": {k: np.asarray(v) + 0.1 for k, v in hp["theta_k"].items()},
           "theta_i": {i: np.asarray(v) - 0.05 for i, v in hp["theta_i"].items()}, "pi_k": hp["pi_k"]}
    l2 = magma.logL_monitoring_VEM(hp2, db, magma.kernel, magma.kernel_mu, got, m_k, session=session)
    r2 = O.logL_monitoring_VEM(hp2, db, O.kernel, O.kernel_mu, ref, m_k)
    assert abs(l2 - r2) <= RTOL_LL * abs(r2)


@pytest.mark.parametrize("common_k,common_i", [(True, True), (True, False), (False, False), (False, True)])
def test_m_step_matches_oracle_optimum(session, common_k, common_i):
    """Optimiser trajectories are not pinned by the reference (third-party L-BFGS-B); the contract is that
    both optimisers reach the same optimum of the same objective: objective values agree and the GPU
    optimum is not worse."""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=10, seed=4, common_hp_i=common_i)
    if common_k:
        first = next(iter(hp["theta_k"].values()))
        hp["theta_k"] = {k: first.copy() for k in hp["theta_k"]}
    ref, got = _setup(session, db, hp, tau, m_k)
    new_ref = O.m_step_VEM(db, hp, ref, O.kernel_mu, O.kernel, m_k, common_k, common_i)
    new_got = magma.m_step_VEM(db, hp, got, magma.kernel_mu, magma.kernel, m_k, common_k, common_i, session=session)
    assert np.allclose(new_got["pi_k"], new_ref["pi_k"], rtol=0, atol=1e-12)
    # individuals' part of the objective at the returned theta_i (CF.R:218-248 summed): GPU optimum not worse
    per_id = O.split_db(db)
    fi_ref = sum(O.logL_clust_multi_GP(new_ref["theta_i"][i], d, ref, O.kernel) for i, d in per_id.items())
    fi_got = sum(O.logL_clust_multi_GP(new_got["theta_i"][i], d, ref, O.kernel) for i, d in per_id.items())
    fi_old = sum(O.logL_clust_multi_GP(hp["theta_i"][i], d, ref, O.kernel) for i, d in per_id.items())
    assert fi_got < fi_old
    assert fi_got <= fi_ref + 1e-6 * abs(fi_ref)
    for i in new_ref["theta_i"]:
        assert np.allclose(new_got["theta_i"][i], new_ref["theta_i"][i], atol=5e-3), i
    # mean processes: (a, b) optimised, third entry = pen_diag = mean_i theta_i[3] (signed)
    for k in new_ref["theta_k"]:
        assert np.allclose(new_got["theta_k"][k], new_ref["theta_k"][k], atol=5e-3), k
    pen = np.mean([new_got["theta_i"][i][2] for i in new_got["theta_i"]])
    assert abs(new_got["theta_k"][next(iter(m_k))][2] - pen) < 1e-12


def test_m_step_with_more_than_eight_clusters(session):
    """K = 9 takes the unfused evaluation of the shared-theta_k objective inside the device-side optimiser loop (the fused
    stages keep the per-cluster products K^-1 r_z in eight registers); same contract as above."""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=27, K=9, seed=6, common_hp_i=False)
    first = next(iter(hp["theta_k"].values()))
    hp["theta_k"] = {k: first.copy() for k in hp["theta_k"]}
    ref, got = _setup(session, db, hp, tau, m_k)
    new_ref = O.m_step_VEM(db, hp, ref, O.kernel_mu, O.kernel, m_k, True, False)
    new_got = magma.m_step_VEM(db, hp, got, magma.kernel_mu, magma.kernel, m_k, True, False, session=session)
    assert np.allclose(new_got["pi_k"], new_ref["pi_k"], rtol=0, atol=1e-12)
    for k in new_ref["theta_k"]:
        assert np.allclose(new_got["theta_k"][k], new_ref["theta_k"][k], atol=5e-3), k
    th_g = next(iter(new_got["theta_k"].values()))
    th_r = next(iter(new_ref["theta_k"].values()))
    # The third entry is pen_diag = mean_i theta_i[3] (CF.R:666-668), i.e. an OUTPUT of the 27 per-individual optimisations,
    # whose end points move in the 7th digit from run to run (atomic scatter in the E-step, DESIGN "Run-to-run
    # reproducibility"): measured over 24 runs |pen_g - pen_r| <= 3e-7, which alone moves this objective by +-1.2e-5
    # relative (gpurun_out/r02s3_flaky.log). So the (a, b) optimum is compared at ONE value of pen_diag, and pen_diag itself
    # separately.
    assert abs(th_g[2] - th_r[2]) < 1e-5
    f_g = O.logL_GP_mod_common_hp_k(th_g[:2], ref["mean"], m_k, O.kernel_mu, ref["cov"], th_r[2])
    f_r = O.logL_GP_mod_common_hp_k(th_r[:2], ref["mean"], m_k, O.kernel_mu, ref["cov"], th_r[2])
    assert f_g <= f_r + 5e-6 * abs(f_r)   # the GPU optimum is not worse (measured: within 2e-7 either way)


def test_per_individual_optimiser_outcomes_are_reported(session):
    """mcb_last_mstep_indiv: optim's counts and convergence code of every per-individual optimisation (CF.R:653-659 keeps only
    `par`): evaluation counts add up to the M-step's total, every optimisation ends in a documented state, and the returned
    theta_i is a point where the oracle's objective is not above its value at the start point"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=16, K=3, n=(8, 20), n_grid=50, seed=41)
    ref, got = _setup(session, db, hp, tau, m_k)
    new = magma.m_step_VEM(db, hp, got, magma.kernel_mu, magma.kernel, m_k, True, False, session=session)
    stats = session.eng.last_mstep_stats()
    nfev, nit, status = session.eng.last_mstep_indiv()
    assert nfev.shape == (16,) and np.all(nfev >= 1) and np.all(nit >= 0) and np.all(nit <= 100)
    assert set(status.tolist()) <= {1, 2, 3, 4}, status          # converged / maxit / abnormal line search / zero gradient
    assert int(nfev.sum()) == stats["nfev_i_sum"] and int(nfev.max()) == stats["nfev_i_max"]
    per = O.split_db(db)
    for i, d in per.items():
        f_new = O.logL_clust_multi_GP(np.asarray(new["theta_i"][i]), d, ref, O.kernel)
        f_old = O.logL_clust_multi_GP(np.asarray(hp["theta_i"][i]), d, ref, O.kernel)
        assert f_new <= f_old + 1e-9 * abs(f_old), i


def test_prediction_into_caller_owned_page_locked_buffers(session):
    """mcb_predict with out = page-locked buffers (mcb_host_alloc) reused across calls, a batch large enough to be cut into
    chunks whose copies overlap the next chunk: identical to the default path"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=12, K=2, n=(8, 14), n_grid=60, seed=43, common_hp_i=True)
    param = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    grid = np.round(np.linspace(0, 10, 60), 3)
    magma.posterior_mu_k(db, grid, m_k, magma.kernel_mu, magma.kernel, {"hp": hp, "param": param}, session=session)
    B, nstar, K = 1300, 6, 2                                   # 1300 / 512 -> 2 chunks
    rng = np.random.default_rng(2)
    obs = np.sort(np.argsort(rng.random((B, 60)), axis=1)[:, :nstar], axis=1).astype(np.int32)
    y = rng.standard_normal((B, nstar))
    off = (np.arange(B + 1) * nstar).astype(np.int32)
    pidx = np.arange(0, 60, 3, dtype=np.int32)
    th = np.array([0.7, 0.4, 0.35])
    e = session.eng
    t0, m0, v0 = e.predict(off, obs.ravel(), y.ravel(), th, pidx)
    out = (e.pinned_empty((B, K)), e.pinned_empty((B, K, pidx.shape[0])), e.pinned_empty((B, K, pidx.shape[0])))
    for _ in range(2):
        out[1][:] = np.nan
        t1, m1, v1 = e.predict(off, obs.ravel(), y.ravel(), th, pidx, out=out)
        assert np.array_equal(t1, t0) and np.array_equal(m1, m0) and np.array_equal(v1, v0)
