"""GPU, BASELINE.json's FULL sizes: the oracle cannot run there in seconds, so parity rests on size-independent properties of
the path: responsibilities are distributions, covariances are symmetric, the hyper-posterior mean is
LINEAR in the outputs for fixed responsibilities (CF.R:690-731), relabelling the clusters permutes the outputs, an inverse
times its matrix is the identity, the predictive mean is affine in the new individual's outputs (MC.R:255-270), and one
M-step does not lower the ELBO (MC.R:70-88). Array-level API (the dict layer of magma.py is not meant for 1e5 individuals)."""
import numpy as np
import pytest

from magmaclust_b200 import synth
from magmaclust_b200.engine import Engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def _start(ds, K):
    M = ds["M"]
    return np.tile([1.0, 1.0, 0.2], (K, 1)), np.tile([1.0, 1.0, 0.2], (M, 1)), np.zeros(K)


def _estep(eng, ds, y=None, tau=None, pi=None, theta_k=None):
    K = ds["K"]
    tk, ti, mk = _start(ds, K)
    if theta_k is not None:
        tk = theta_k
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"] if y is None else y, ds["grid_full"])
    t0 = ds["tau0"] if tau is None else tau
    eng.set_state(tk, ti, t0.mean(0) if pi is None else pi, mk, t0)
    eng.e_step()
    return eng.get_mean(), eng.get_cov(), eng.get_tau_new()


def test_c2_full_size_properties(eng):
    """C2: M = 1000 individuals x 50 observations, K = 5, grid of 401 points, individual-specific hyper-parameters"""
    c = synth.CONFIGS["C2"]
    ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
    mean, cov, tau = _estep(eng, ds)
    assert np.all(np.isfinite(mean)) and np.all(np.isfinite(cov))
    assert np.abs(tau.sum(1) - 1.0).max() < 1e-12 and tau.min() >= 0.0           # responsibilities are distributions
    assert np.abs(cov - cov.transpose(0, 2, 1)).max() == 0.0                        # exactly symmetric
    assert all(np.diag(cov[k]).min() > 0 for k in range(c["K"]))
    # idempotence: the same call again (the scatter uses atomics: equal up to summation order)
    mean2, cov2, tau2 = _estep(eng, ds)
    assert np.abs(mean2 - mean).max() <= 1e-10 * np.abs(mean).max()
    assert np.abs(tau2 - tau).max() < 1e-9
    # linearity in the outputs for fixed (old) responsibilities: y -> 2 y doubles the means, leaves the covariances
    mean3, cov3, _ = _estep(eng, ds, y=2.0 * ds["y"])
    assert np.abs(mean3 - 2.0 * mean).max() <= 1e-10 * np.abs(mean).max()
    assert np.abs(cov3 - cov).max() <= 1e-10 * np.abs(cov).max()
    # relabelling the clusters permutes everything
    perm = np.array([3, 0, 4, 1, 2])
    mean4, cov4, tau4 = _estep(eng, ds, tau=ds["tau0"][:, perm], pi=ds["tau0"].mean(0)[perm])
    assert np.abs(mean4 - mean[perm]).max() <= 1e-10 * np.abs(mean).max()
    assert np.abs(tau4 - tau[:, perm]).max() < 1e-9
    assert np.array_equal(tau4.argmax(1), np.argsort(perm)[tau.argmax(1)]) or np.array_equal(perm[tau4.argmax(1)], tau.argmax(1))
    # one VEM iteration from the study's start state: the M-step does not lower the bound (MC.R:70: eps > 0)
    tk, ti, mk = _start(ds, c["K"])
    eng.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
    elbo, eps, conv = eng.vem_iteration(True, False)
    assert np.isfinite(elbo) and eps > 0.0
    new_tk, new_ti, new_pi = eng.get_new_hp()
    assert abs(new_pi.sum() - 1.0) < 1e-12 and np.all(np.isfinite(new_ti))
    assert np.allclose(new_tk[:, 2], new_ti[:, 2].mean())                           # pen_diag = mean_i theta_i[3], CF.R:662


def test_c3_full_size_properties(eng):
    """C3: M = 100,000 individuals x 30 observations, K = 5, grid of 201 points, shared hyper-parameters"""
    c = synth.CONFIGS["C3"]
    ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"])
    mean, cov, tau = _estep(eng, ds)
    L = eng.get_mat_logL()
    assert np.all(np.isfinite(mean)) and np.all(np.isfinite(cov))
    assert np.abs(tau.sum(1) - 1.0).max() < 1e-12 and tau.min() >= 0.0
    assert np.abs(cov - cov.transpose(0, 2, 1)).max() == 0.0
    mean3, cov3, _ = _estep(eng, ds, y=-0.5 * ds["y"])
    assert np.abs(mean3 + 0.5 * mean).max() <= 1e-9 * np.abs(mean).max()
    assert np.abs(cov3 - cov).max() <= 1e-9 * np.abs(cov).max()
    # the whole 1e5 x 5 table of responsibilities against the update rule applied on the host to the device's own
    # log-likelihoods (max-shift over the clusters, pi multiplied afterwards, CF.R:755-757)
    assert np.all(np.isfinite(L))
    w = ds["tau0"].mean(0)[None, :] * np.exp(L - L.max(1, keepdims=True))
    w /= w.sum(1, keepdims=True)
    assert np.abs(w - tau).max() < 1e-12
    # shared theta_i: objective and gradient of the summed individual likelihood, against central differences (f64)
    th = np.array([0.9, 0.7, 0.3])
    f0, g0 = eng.eval_indiv_common(th)
    for j in range(3):
        h = 1e-5
        e = np.zeros(3)
        e[j] = h
        fp, _ = eng.eval_indiv_common(th + e)
        fm, _ = eng.eval_indiv_common(th - e)
        assert abs((fp - fm) / (2 * h) - g0[j]) <= 1e-5 * max(1.0, abs(g0[j]))


def test_c4_full_size_inverse(eng):
    """C4: the N = 8192 hyper-posterior solve. K^-1 from the large-grid path times K is the identity (random probes), the
    inverse is exactly symmetric, log det agrees with the sum over a split of the spectrum bound"""
    N = 8192
    x = np.linspace(0.0, 10.0, N)
    theta2, sigma = np.array([1.0, 1.0]), 0.2
    Kinv, logdet = eng.kern_to_inv(x, theta2, sigma)
    assert np.abs(Kinv - Kinv.T).max() == 0.0
    d = x[:, None] - x[None, :]
    Kmat = np.exp(theta2[0] - 0.5 * np.exp(-theta2[1]) * d * d)
    Kmat[np.diag_indices(N)] = np.exp(theta2[0]) + sigma ** 2
    rng = np.random.default_rng(0)
    V = rng.standard_normal((N, 4))
    R = Kmat @ (Kinv @ V) - V
    assert np.abs(R).max() <= 1e-8 * np.abs(V).max()
    # log det: bounded by N log(largest eigenvalue bound) and N log(sigma^2); and it must reproduce on a second call
    assert N * np.log(sigma ** 2) < logdet < N * np.log(N * np.exp(theta2[0]) + sigma ** 2)
    _, logdet2 = eng.kern_to_inv(x, theta2, sigma)
    assert logdet2 == logdet


def test_c5_full_size_prediction_properties(eng):
    """C5: 10,000 new individuals x 20 observations onto a 2,000-point grid, K = 5"""
    B, T, K, nstar = 10000, 2000, 5, 20
    ds = synth.make_dataset(M=1000, K=K, n=30, n_grid=T, common_hp_i=True, seed=1005)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
    tk, ti, mk = _start(ds, K)
    eng.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
    eng.posterior_mu_k(ds["grid_full"], tau=ds["tau0"], want_mean=False, want_cov=False)
    rng = np.random.default_rng(7)
    obs = np.sort(np.argsort(rng.random((B, T)), axis=1)[:, :nstar], axis=1).astype(np.int32)
    mu = ds["mu_true"][rng.integers(0, K, B)]
    y1 = np.take_along_axis(mu, obs, axis=1) + 0.3 * rng.standard_normal((B, nstar))
    y2 = y1 + rng.standard_normal((B, nstar))
    offsets = (np.arange(B + 1) * nstar).astype(np.int32)
    pred_idx = np.arange(0, T, 8, dtype=np.int32)      # 250 targets: keeps the three result tables at 100 MB each
    theta_new = np.array([1.0, 1.0, 0.2])
    tau1, m1, v1 = eng.predict(offsets, obs.ravel(), y1.ravel(), theta_new, pred_idx)
    tau2, m2, v2 = eng.predict(offsets, obs.ravel(), y2.ravel(), theta_new, pred_idx)
    tau3, m3, v3 = eng.predict(offsets, obs.ravel(), (0.5 * (y1 + y2)).ravel(), theta_new, pred_idx)
    for t in (tau1, tau2, tau3):
        assert np.abs(t.sum(1) - 1.0).max() < 1e-12 and t.min() >= 0.0
    assert v1.min() > 0.0 and np.array_equal(v1, v2) and np.array_equal(v1, v3)    # the variance does not depend on y
    # per cluster the predictive mean is affine in y: the midpoint of the inputs gives the midpoint of the outputs
    assert np.abs(m3 - 0.5 * (m1 + m2)).max() <= 1e-9 * max(np.abs(m1).max(), np.abs(m2).max())


@pytest.mark.parametrize("common_i", [False, True])
def test_vem_iteration_host_delivers_the_e_step_results(eng, common_i):
    """mcb_vem_iteration_host: the copies that leave behind the M-step carry exactly what the getters return, and the
    iteration's statistics are those of the plain call from the same state"""
    ds = synth.make_dataset(M=300, K=3, n=24, n_grid=97, common_hp_i=common_i, seed=77)
    K, M, N = 3, 300, 97
    tk, ti, mk = _start(ds, K)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
    eng.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
    ref = eng.vem_iteration(True, common_i)
    ref_hp = eng.get_new_hp()
    ref_mean, ref_cov, ref_tau = eng.get_mean(), eng.get_cov(), eng.get_tau_new()
    eng.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
    om, oc, ot = np.full((K, N), np.nan), np.full((K, N, N), np.nan), np.full((M, K), np.nan)
    got = eng.vem_iteration_host(True, common_i, om, oc, ot)
    assert np.array_equal(om, eng.get_mean()) and np.array_equal(oc, eng.get_cov()) and np.array_equal(ot, eng.get_tau_new())
    assert np.abs(om - ref_mean).max() <= 1e-10 * np.abs(ref_mean).max()       # scatter order differs between calls
    assert np.abs(oc - ref_cov).max() <= 1e-10 * np.abs(ref_cov).max()
    assert np.abs(ot - ref_tau).max() < 1e-9
    assert abs(got[0] - ref[0]) <= 1e-8 * abs(ref[0]) and got[2] == ref[2]
    # the optimisers stop on a relative decrease of the objective: two runs whose inputs differ in the last bits agree in
    # the bound (above) far more tightly than in the position along its flat directions
    for a, b in zip(eng.get_new_hp(), ref_hp):
        assert np.all(np.isfinite(a)) and np.abs(a - b).max() <= 5e-2 * max(1.0, np.abs(b).max())
    # any subset of the three destinations may be absent
    eng.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
    ot2 = np.empty((M, K))
    eng.vem_iteration_host(True, common_i, None, None, ot2)
    assert np.abs(ot2 - ot).max() < 1e-9


def test_same_shape_dataset_reupload_keeps_nothing_of_the_old_data(eng):
    """mcb_set_data with a dataset of the SAME shape keeps the state buffers and the captured theta_k graphs (the host-buffer
    loop relies on it): the iteration on the second dataset must equal the one of a fresh handle that never saw the first"""
    K = 3
    a = synth.make_dataset(M=200, K=K, n=20, n_grid=81, common_hp_i=False, seed=5)
    b = synth.make_dataset(M=200, K=K, n=20, n_grid=81, common_hp_i=False, seed=6)
    assert a["y"].shape == b["y"].shape and not np.array_equal(a["y"], b["y"])
    tk, ti, mk = _start(a, K)
    for ds in (a, b):
        eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
        eng.set_state(tk, ti, ds["tau0"].mean(0), mk, ds["tau0"])
        got = eng.vem_iteration(True, False)
    got_mean, got_tau = eng.get_mean(), eng.get_tau_new()
    fresh = Engine(0)
    try:
        fresh.set_data(b["offsets"], b["grid_idx_full"], b["y"], b["grid_full"])
        fresh.set_state(tk, ti, b["tau0"].mean(0), mk, b["tau0"])
        ref = fresh.vem_iteration(True, False)
        assert np.abs(got_mean - fresh.get_mean()).max() <= 1e-10 * np.abs(got_mean).max()
        assert np.abs(got_tau - fresh.get_tau_new()).max() < 1e-9
        assert abs(got[0] - ref[0]) <= 1e-8 * abs(ref[0])
    finally:
        fresh.close()
