"""GPU vs oracle AT THE BASELINE.json CONFIGURATIONS (SURVEY.md §8a sizes C1..C5), tolerances of north_star:
1e-8 relative on hyper-posterior means / covariances, 1e-6 on log-likelihoods / responsibilities, identical hard assignments.

  C1  full training_VEM against the oracle's run (M = 50, K = 3, n = 30, common hp)
  C2  full e_step_VEM against the oracle (M = 1000, n = 50, N = 401), logL_clust_multi_GP / gr_clust_multi_GP on 50 sampled
      individuals, logL_GP_mod_common_hp_k / gr_GP_mod_common_hp_k at N = 401, the ELBO monitor
  C3  M = 100,000: the oracle cannot run the E-step in seconds, so 200 sampled individuals' responsibilities, corr terms and
      objective / gradient values are recomputed by the oracle FROM THE GPU'S OWN (mean, cov) (CF.R:733-762, 218-248), and the
      hyper-posterior itself is checked by the defining identity A_k C_k = I, A_k built on the host from oracle pieces
  C4  N = 8192, K = 8 (M = 4096): full E-step; every cluster's C_k against the defining identity (random probes, without
      inverting anything on the host), sampled responsibilities against the oracle
  C5  10,000 new individuals on a 2,000-point grid: sampled individuals' tau* / Mean / Var against O.update_tau_star_k_EM /
      O.pred_gp_clust fed the GPU's prediction posterior (CF.R:764-786, MC.R:235-274)
"""
import numpy as np
import pytest

from magmaclust_b200 import synth
from magmaclust_b200.engine import Engine
from oracle import magma_oracle as O
from helpers import RTOL_LL, RTOL_POST, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def _names(K):
    return [f"K{k + 1}" for k in range(K)]


def _start(M, K):
    return np.tile([1.0, 1.0, 0.2], (K, 1)), np.tile([1.0, 1.0, 0.2], (M, 1)), np.zeros(K)


def _sub_db(ds, rows):
    """R-style db of the individuals `rows` (positions) of a synthetic dataset"""
    o = ds["offsets"]
    sel = np.concatenate([np.arange(o[r], o[r + 1]) for r in rows])
    n = np.diff(o)[rows]
    return {"ID": np.repeat(ds["ids"][rows], n), "Timestamp": ds["t"][sel], "Output": ds["y"][sel]}


def _param(grid, mean, cov, tau=None, ids=None):
    names = _names(mean.shape[0])
    p = {"mean": {k: {"Timestamp": grid, "Output": mean[c]} for c, k in enumerate(names)},
         "cov": {k: O.NamedMat(cov[c], grid) for c, k in enumerate(names)}}
    if tau is not None:
        p["tau_i_k"] = {k: {int(i): float(tau[r, c]) for r, i in enumerate(ids)} for c, k in enumerate(names)}
    return p


# ---------------------------------------------------------------------------------------------------------------------
def test_c1_full_training_matches_the_oracle(session):
    """C1 (Simulation_study.R's synthetic set: M = 50, K = 3, ~30 observations on [0, 10], SE kernel, common hp): the whole
    training_VEM run, GPU against oracle: same number of iterations, ELBO trace, responsibilities, hyper-posterior"""
    from magmaclust_b200 import magma
    c = synth.CONFIGS["C1"]
    ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
    db = synth.to_db(ds)
    names = _names(c["K"])
    tau0 = synth.tau_dict(ds, names)
    m_k = {k: 0.0 for k in names}
    ini = {"theta_k": np.array([1.0, 1.0, 0.2]), "theta_i": np.array([1.0, 1.0, 0.2])}
    ref = O.training_VEM(db, m_k, ini, O.kernel_mu, O.kernel, tau0, True, True)
    got = magma.training_VEM(db, m_k, ini, magma.kernel_mu, magma.kernel, tau0, True, True, session=session)
    assert got["convergence"] == ref["convergence"]
    assert len(got["logL_iterations"]) == len(ref["logL_iterations"])
    # the ELBO crosses zero along the run (-6104 ... +272): 1e-4 of its range, not of each (possibly tiny) value. Later
    # iterations inherit the end points of two different optimisers (library L-BFGS vs scipy), which agree to optimiser
    # tolerance only
    assert np.abs(np.asarray(got["logL_iterations"]) - np.asarray(ref["logL_iterations"])).max() <= \
        1e-4 * np.abs(ref["logL_iterations"]).max()
    assert np.array_equal(magma.pred_max_cluster(got["param"]["tau_i_k"]), O.pred_max_cluster(ref["param"]["tau_i_k"]))
    # first iteration: same inputs on both sides -> the contract tolerances apply to its E-step (later iterations inherit the
    # two optimisers' end points, which agree to optimiser tolerance only)
    hp0 = {"theta_k": {k: ini["theta_k"] for k in names}, "theta_i": {int(i): ini["theta_i"] for i in ds["ids"]},
           "pi_k": np.array([np.mean(list(tau0[k].values())) for k in tau0])}
    r0 = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp0, tau0)
    g0 = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp0, tau0, session=session)
    for k in names:
        assert rel_err(g0["mean"][k]["Output"], r0["mean"][k]["Output"]) < RTOL_POST
        assert rel_err(g0["cov"][k].m, r0["cov"][k].m) < RTOL_POST
    A, B = O.tau_to_array(g0["tau_i_k"]), O.tau_to_array(r0["tau_i_k"])
    assert np.abs(A - B).max() < RTOL_LL and np.array_equal(A.argmax(1), B.argmax(1))
    for k in names:
        assert np.allclose(got["hp"]["theta_k"][k], ref["hp"]["theta_k"][k], atol=2e-2)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2():
    c = synth.CONFIGS["C2"]
    ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
    db = synth.to_db(ds)
    names = _names(c["K"])
    rng = np.random.default_rng(11)
    M = c["M"]
    # individual-specific hyper-parameters away from the start state, soft responsibilities
    th = np.stack([rng.uniform(0.5, 2.0, M), rng.uniform(0.3, 1.5, M), rng.uniform(0.15, 0.5, M)], axis=1)
    w = 0.9 * ds["tau0"] + 0.1 * rng.dirichlet(np.ones(c["K"]), M)
    hp = {"theta_k": {k: np.array([1.3, 0.9, 0.25]) for k in names},
          "theta_i": {int(i): th[r] for r, i in enumerate(ds["ids"])}, "pi_k": w.mean(0)}
    tau = {k: {int(i): float(w[r, cc]) for r, i in enumerate(ds["ids"])} for cc, k in enumerate(names)}
    m_k = {k: 0.5 * cc for cc, k in enumerate(names)}
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau, return_logL=True)
    return dict(ds=ds, db=db, names=names, hp=hp, tau=tau, m_k=m_k, ref=ref, th=th)


def test_c2_full_e_step_matches_the_oracle(session, c2):
    from magmaclust_b200 import magma
    got = magma.e_step_VEM(c2["db"], c2["m_k"], magma.kernel_mu, magma.kernel, c2["hp"], c2["tau"], session=session)
    ref = c2["ref"]
    for k in c2["names"]:
        assert rel_err(got["mean"][k]["Output"], ref["mean"][k]["Output"]) < RTOL_POST, k
        assert rel_err(got["cov"][k].m, ref["cov"][k].m) < RTOL_POST, k
    A, B = O.tau_to_array(got["tau_i_k"]), O.tau_to_array(ref["tau_i_k"])
    assert np.abs(A - B).max() < RTOL_LL
    assert np.array_equal(A.argmax(1), B.argmax(1))
    L = got["mat_logL"]
    assert np.abs(L - ref["mat_logL"]).max() <= RTOL_LL * np.abs(ref["mat_logL"]).max()


def test_c2_m_step_callbacks_match_the_oracle(session, c2):
    """the optimiser callbacks at C2 size: per-individual objective / gradient on 50 sampled individuals, the common-theta_k
    objective / gradient at N = 401 (resident AND explicit-argument forms), the ELBO monitor"""
    from magmaclust_b200 import magma
    s = session
    got = magma.e_step_VEM(c2["db"], c2["m_k"], magma.kernel_mu, magma.kernel, c2["hp"], c2["tau"], session=s)
    ref = c2["ref"]
    rng = np.random.default_rng(3)
    M = c2["ds"]["M"]
    trial = c2["th"] + rng.normal(0, 0.1, (M, 3))
    f, g = s.eng.eval_indiv(trial)
    per = O.split_db(c2["db"])
    ids = [int(i) for i in c2["ds"]["ids"]]
    for r in rng.choice(M, 50, replace=False):
        fr = O.logL_clust_multi_GP(trial[r], per[ids[r]], ref, O.kernel)
        gr = O.gr_clust_multi_GP(trial[r], per[ids[r]], ref, O.kernel)
        assert abs(f[r] - fr) <= RTOL_LL * max(1.0, abs(fr)), r
        assert np.abs(g[r] - gr).max() <= RTOL_LL * max(1.0, np.abs(gr).max()), r
    th3 = np.array([1.1, 0.7, 0.3])
    for nhp in (3, 2):
        fr = O.logL_GP_mod_common_hp_k(th3[:nhp], ref["mean"], c2["m_k"], O.kernel_mu, ref["cov"], 0.27)
        gr = O.gr_GP_mod_common_hp_k(th3[:nhp], ref["mean"], c2["m_k"], O.kernel_mu, ref["cov"], 0.27)
        # resident form: the very objects e_step_VEM returned
        fg = magma.logL_GP_mod_common_hp_k(th3[:nhp], got["mean"], c2["m_k"], magma.kernel_mu, got["cov"], 0.27, session=s)
        gg = magma.gr_GP_mod_common_hp_k(th3[:nhp], got["mean"], c2["m_k"], magma.kernel_mu, got["cov"], 0.27, session=s)
        assert abs(fg - fr) <= RTOL_LL * abs(fr) and np.abs(gg - gr).max() <= RTOL_LL * max(1.0, np.abs(gr).max())
        # explicit form: the oracle's objects (not resident) are uploaded
        fe = magma.logL_GP_mod_common_hp_k(th3[:nhp], ref["mean"], c2["m_k"], magma.kernel_mu, ref["cov"], 0.27, session=s)
        ge = magma.gr_GP_mod_common_hp_k(th3[:nhp], ref["mean"], c2["m_k"], magma.kernel_mu, ref["cov"], 0.27, session=s)
        assert abs(fe - fr) <= RTOL_LL * abs(fr) and np.abs(ge - gr).max() <= RTOL_LL * max(1.0, np.abs(gr).max())
    k = c2["names"][2]
    fr = O.logL_GP_mod(th3, ref["mean"][k], c2["m_k"][k], O.kernel_mu, ref["cov"][k])
    gr = O.gr_GP_mod(th3, ref["mean"][k], c2["m_k"][k], O.kernel_mu, ref["cov"][k])
    for mean_arg, cov_arg in ((got["mean"][k], got["cov"][k]), (ref["mean"][k], ref["cov"][k])):
        fg = magma.logL_GP_mod(th3, mean_arg, c2["m_k"][k], magma.kernel_mu, cov_arg, session=s)
        gg = magma.gr_GP_mod(th3, mean_arg, c2["m_k"][k], magma.kernel_mu, cov_arg, session=s)
        assert abs(fg - fr) <= RTOL_LL * abs(fr) and np.abs(gg - gr).max() <= RTOL_LL * max(1.0, np.abs(gr).max())
    e_ref = O.logL_monitoring_VEM(c2["hp"], c2["db"], O.kernel, O.kernel_mu, ref, c2["m_k"])
    e_got = magma.logL_monitoring_VEM(c2["hp"], c2["db"], magma.kernel, magma.kernel_mu, got, c2["m_k"], session=s)
    assert abs(e_got - e_ref) <= RTOL_LL * abs(e_ref)


# ---------------------------------------------------------------------------------------------------------------------
def _batch_inv(t, theta):
    """Psi_i^-1 for a batch of individuals with n observations each: kern_to_cov (CF.R:61-86) + solve (LU, CF.R:142), the
    oracle's arithmetic written for a [B, n] block of timestamps. theta: [B, 3]"""
    d = t[:, :, None] - t[:, None, :]
    Psi = np.exp(theta[:, 0, None, None] - 0.5 * np.exp(-theta[:, 1, None, None]) * d * d)
    n = t.shape[1]
    Psi[:, np.arange(n), np.arange(n)] = (np.exp(theta[:, 0]) + theta[:, 2] ** 2)[:, None]
    return np.linalg.inv(Psi)


def _host_accumulators(ds, theta_i, tau, N):
    """S_k = sum_i tau_ik P_i' Psi_i^-1 P_i and v_k = sum_i tau_ik P_i' Psi_i^-1 y_i for every cluster (CF.R:698-705, 719-729),
    dense N x N / N, from oracle arithmetic"""
    M, n, K = ds["M"], ds["n"], tau.shape[1]
    gi = ds["grid_idx_full"].reshape(M, n).astype(np.int64)
    t = ds["t"].reshape(M, n)
    y = ds["y"].reshape(M, n)
    S = np.zeros((K, N * N))
    v = np.zeros((K, N))
    for s0 in range(0, M, 8192):
        s1 = min(M, s0 + 8192)
        W = _batch_inv(t[s0:s1], theta_i[s0:s1])
        flat = (gi[s0:s1, :, None] * N + gi[s0:s1, None, :]).ravel()
        Wy = np.einsum("bij,bj->bi", W, y[s0:s1])
        for k in range(K):
            wk = tau[s0:s1, k]
            S[k] += np.bincount(flat, weights=(wk[:, None, None] * W).ravel(), minlength=N * N)
            v[k] += np.bincount(gi[s0:s1].ravel(), weights=(wk[:, None] * Wy).ravel(), minlength=N)
    return S.reshape(K, N, N), v


def test_c3_sampled_individuals_match_the_oracle(eng):
    """C3: M = 100,000, K = 5, n_i = 30, N = 201, shared hyper-parameters"""
    c = synth.CONFIGS["C3"]
    ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
    M, K = c["M"], c["K"]
    grid = ds["grid_full"]
    rng = np.random.default_rng(5)
    w = 0.95 * ds["tau0"] + 0.05 * rng.dirichlet(np.ones(K), M)
    tk = np.tile([1.2, 0.8, 0.3], (K, 1))
    ti = np.tile([0.9, 0.6, 0.35], (M, 1))
    mk = np.zeros(K)
    pi = w.mean(0)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], grid)
    eng.set_state(tk, ti, pi, mk, w)
    eng.e_step()
    mean, cov, tau, L = eng.get_mean(), eng.get_cov(), eng.get_tau_new(), eng.get_mat_logL()
    names = _names(K)
    # (1) responsibilities / log-likelihoods of 200 sampled individuals from the GPU's own hyper-posterior
    rows = np.sort(rng.choice(M, 200, replace=False))
    db_s = _sub_db(ds, rows)
    p = _param(grid, mean, cov)
    hp_s = {"theta_i": {int(ds["ids"][r]): ti[r] for r in rows}}
    tau_ref, L_ref = O.update_tau_i_k_VEM(db_s, {k: 0.0 for k in names}, p["mean"], p["cov"], O.kernel, hp_s, pi)
    T_ref = O.tau_to_array(tau_ref)
    assert np.abs(tau[rows] - T_ref).max() < RTOL_LL
    assert np.array_equal(tau[rows].argmax(1), T_ref.argmax(1))
    assert np.abs(L[rows] - L_ref.T).max() <= RTOL_LL * np.abs(L_ref).max()
    # (2) the shared-theta_i objective / gradient restricted to the sample: per-individual values of the batched evaluator
    p["tau_i_k"] = {k: {int(ds["ids"][r]): float(tau[r, cc]) for r in rows} for cc, k in enumerate(names)}
    trial = np.tile([1.1, 0.5, 0.3], (M, 1))
    f, g = eng.eval_indiv(trial)
    per = O.split_db(db_s)
    for r in rows[:60]:
        i = int(ds["ids"][r])
        fr = O.logL_clust_multi_GP(trial[r], per[i], p, O.kernel)
        gr = O.gr_clust_multi_GP(trial[r], per[i], p, O.kernel)
        assert abs(f[r] - fr) <= RTOL_LL * max(1.0, abs(fr))
        assert np.abs(g[r] - gr).max() <= RTOL_LL * max(1.0, np.abs(gr).max())
    # and the summed form the optimiser sees (all 1e5 individuals, CF.R:280-325 / 539-586) equals the sum of those values
    fs, gs = eng.eval_indiv_common(trial[0])
    assert abs(fs - f.sum()) <= 1e-10 * abs(fs) and np.abs(gs - g.sum(0)).max() <= 1e-9 * np.abs(gs).max()
    # (3) the hyper-posterior itself against the oracle's arithmetic on ALL 1e5 individuals: A_k = K^-1 + S_k assembled on the
    # host, C_k = solve(A_k) (CF.R:612), m_hat_k = C_k (K^-1 m_k + v_k) (CF.R:620) -- the 1e-8 contract on means / covariances
    N = grid.shape[0]
    Kinv = O.kern_to_inv(grid, O.kernel_mu, tk[0, :2], tk[0, 2]).m
    S, v = _host_accumulators(ds, ti, w, N)
    for k in range(K):
        C_ref = np.linalg.inv(Kinv + S[k])
        assert rel_err(cov[k], C_ref) < RTOL_POST, k
        assert rel_err(mean[k], C_ref @ v[k]) < RTOL_POST, k


# ---------------------------------------------------------------------------------------------------------------------
def test_c4_full_e_step_hyperposterior_identity_and_sampled_tau(eng):
    """C4: dense union grid N = 8192, K = 8, M = 4096 individuals x 32 observations (every grid point touched). The full
    E-step; every C_k against A_k C_k = I, rewritten as  C_k + K (S_k C_k) = K  so that nothing is inverted on the host
    (K = prior covariance, S_k = sum_i tau_ik P_i' Psi_i^-1 P_i), on random probe vectors; sampled responsibilities"""
    c = synth.CONFIGS["C4"]
    ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"], equispaced_cover=True)
    M, K, N = c["M"], c["K"], c["n_grid"]
    grid = ds["grid_full"]
    tk, ti, mk = _start(M, K)
    tau0 = ds["tau0"]
    pi = tau0.mean(0)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], grid)
    eng.set_state(tk, ti, pi, mk, tau0)
    eng.e_step()
    mean, tau = eng.get_mean(), eng.get_tau_new()
    assert np.abs(tau.sum(1) - 1.0).max() < 1e-12
    d = grid[:, None] - grid[None, :]
    Kmat = np.exp(tk[0, 0] - 0.5 * np.exp(-tk[0, 1]) * d * d)
    Kmat[np.diag_indices(N)] = np.exp(tk[0, 0]) + tk[0, 2] ** 2
    del d
    rng = np.random.default_rng(2)
    V = rng.standard_normal((N, 3))
    KV = Kmat @ V
    o, gi, t = ds["offsets"], ds["grid_idx_full"], ds["t"]
    n = c["n"]
    W = _batch_inv(t.reshape(M, n), ti)
    G2 = gi.reshape(M, n).astype(np.int64)
    Y2 = ds["y"].reshape(M, n)
    Wy = np.einsum("bij,bj->bi", W, Y2)
    rows = np.sort(rng.choice(M, 40, replace=False))
    names = _names(K)
    covs = {}
    Kinv_host = None
    for k in range(K):
        Ck = eng.get_cov(k)
        assert np.abs(Ck - Ck.T).max() == 0.0
        CV = Ck @ V
        mem = np.nonzero(tau0[:, k] > 0)[0]
        # S_k (C_k V): per member individual  P_i' (tau W_i) P_i  applied to the rows idx_i of C_k V
        contrib = np.einsum("bij,bjm->bim", tau0[mem, k, None, None] * W[mem], CV[G2[mem]])
        SCV = np.zeros_like(CV)
        np.add.at(SCV, G2[mem].ravel(), contrib.reshape(-1, CV.shape[1]))
        R = CV + Kmat @ SCV - KV
        assert np.abs(R).max() <= 1e-7 * np.abs(KV).max(), (k, np.abs(R).max())
        wk = np.bincount(G2[mem].ravel(), weights=(tau0[mem, k, None] * Wy[mem]).ravel(), minlength=N)
        assert np.abs(Ck @ wk - mean[k]).max() <= 1e-8 * np.abs(mean[k]).max()
        if k in (0, 5):
            # forward comparison against the oracle's arithmetic (LU solves on the host): C_k V = solve(K^-1 + S_k, V)
            if Kinv_host is None:
                Kinv_host = np.linalg.inv(Kmat)
            flat = (G2[mem][:, :, None] * N + G2[mem][:, None, :]).ravel()
            A = Kinv_host + np.bincount(flat, weights=(tau0[mem, k, None, None] * W[mem]).ravel(),
                                        minlength=N * N).reshape(N, N)
            X = np.linalg.solve(A, np.c_[V, wk])
            del A
            assert rel_err(CV, X[:, :3]) < RTOL_POST, k
            assert rel_err(mean[k], X[:, 3]) < RTOL_POST, k
        # keep only what the sampled individuals need: C_k[idx_i, idx_i]
        covs[k] = {int(r): Ck[np.ix_(gi[o[r]:o[r + 1]], gi[o[r]:o[r + 1]])].copy() for r in rows}
        del Ck
    # responsibilities of the sampled individuals: l_ik = -logL_GP_mod(theta_i, db_i, m_hat_k[idx], C_k[idx, idx]) (CF.R:746)
    for r in rows:
        idx = gi[o[r]:o[r + 1]]
        dbi = {"Timestamp": t[o[r]:o[r + 1]], "Output": ds["y"][o[r]:o[r + 1]]}
        l = np.array([-O.logL_GP_mod(ti[r], dbi, mean[k][idx], O.kernel, covs[k][int(r)]) for k in range(K)])
        wgt = pi * np.exp(l - l.max())
        wgt /= wgt.sum()
        assert np.abs(tau[r] - wgt).max() < RTOL_LL
        assert tau[r].argmax() == wgt.argmax()


# ---------------------------------------------------------------------------------------------------------------------
def test_c5_sampled_predictions_match_the_oracle(eng):
    """C5: pred_magmaclust-style prediction, 10,000 new individuals x 20 observations onto a 2,000-point grid, K = 5"""
    B, T, K, nstar = 10000, 2000, 5, 20
    c = synth.CONFIGS["C5"]
    ds = synth.make_dataset(c["M"], K, c["n"], T, c["common_hp_i"], c["seed"])
    grid = ds["grid_full"]
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], grid)
    tk, ti, mk = _start(ds["M"], K)
    eng.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
    mean, cov = eng.posterior_mu_k(grid, tau=ds["tau0"])
    rng = np.random.default_rng(7)
    obs = np.sort(np.argsort(rng.random((B, T)), axis=1)[:, :nstar], axis=1).astype(np.int32)
    mu = ds["mu_true"][rng.integers(0, K, B)]
    y = np.take_along_axis(mu, obs, axis=1) + 0.3 * rng.standard_normal((B, nstar))
    offsets = (np.arange(B + 1) * nstar).astype(np.int32)
    pred_idx = np.arange(0, T, 5, dtype=np.int32)       # 400 targets
    theta_new = np.array([1.0, 1.0, 0.2])
    tau, pm, pv = eng.predict(offsets, obs.ravel(), y.ravel(), theta_new, pred_idx)
    names = _names(K)
    p = _param(grid, mean, cov)
    pi = {k: float(ds["tau0"][:, cc].mean()) for cc, k in enumerate(names)}
    for b in rng.choice(B, 24, replace=False):
        dbn = {"ID": np.full(nstar, 1), "Timestamp": grid[obs[b]], "Output": y[b]}
        t_ref = O.update_tau_star_k_EM(dbn, p["mean"], p["cov"], O.kernel, theta_new, pi)
        assert max(abs(tau[b, cc] - t_ref[k]) for cc, k in enumerate(names)) < RTOL_LL
        pr = O.pred_gp_clust(dbn, grid[pred_idx], p, O.kernel, {"theta_new": theta_new, "tau_k": t_ref})
        for cc, k in enumerate(names):
            assert rel_err(pm[b, cc], pr[k]["Mean"]) < RTOL_POST, (b, k)
            assert np.abs(pv[b, cc] - pr[k]["Var"]).max() <= RTOL_POST * np.abs(pr[k]["Var"]).max() + 1e-12, (b, k)


# ---------------------------------------------------------------------------------------------------------------------
def test_training_then_posterior_on_a_larger_grid_then_training_again(eng):
    """the captured theta_k graphs hold the workspace pointers of the dense inverse. posterior_mu_k on a
    grid with more 32-row blocks than the training grid used to grow that workspace; training the next dataset of the same
    shape then replayed graphs with freed buffers. Now: own workspace for everything outside training. Compare with a
    fresh handle that never ran posterior_mu_k."""
    K = 3
    a = synth.make_dataset(M=150, K=K, n=20, n_grid=81, common_hp_i=False, seed=15)
    b = synth.make_dataset(M=150, K=K, n=20, n_grid=81, common_hp_i=False, seed=16)
    tk, ti, mk = _start(150, K)

    def train(e, ds):
        e.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
        e.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
        return e.vem_iteration(True, False)

    train(eng, a)
    big = np.unique(np.r_[a["grid_full"], np.linspace(0.0, 10.0, 801)])     # 26 blocks of 32 against 3
    eng.posterior_mu_k(big, tau=a["tau0"], want_mean=False, want_cov=False)
    eng.kern_to_inv(np.linspace(0, 10, 500), np.array([1.0, 1.0]), 0.3)
    got = train(eng, b)
    got_mean, got_tau, got_hp = eng.get_mean(), eng.get_tau_new(), eng.get_new_hp()
    fresh = Engine(0)
    try:
        ref = train(fresh, b)
        assert np.abs(got_mean - fresh.get_mean()).max() <= 1e-10 * np.abs(got_mean).max()
        assert np.abs(got_tau - fresh.get_tau_new()).max() < 1e-9
        # eps goes through 151 optimisations: the E-step's atomic scatter makes their inputs differ in the last bit from run
        # to run, and an optimisation in a flat valley can then stop elsewhere (measured between two FRESH handles:
        # d theta_i up to 8e-5, once 0.18 for one individual, eps 3.9725 vs 3.9508; tools/dbg/reuse_check.py). A replayed graph
        # on freed buffers gives garbage or a fault, not a 0.5 % shift.
        assert abs(got[0] - ref[0]) <= 1e-8 * abs(ref[0]) and abs(got[1] - ref[1]) <= 2e-2 * abs(ref[1])
        assert np.abs(got_hp[0] - fresh.get_new_hp()[0]).max() < 5e-2
    finally:
        fresh.close()
