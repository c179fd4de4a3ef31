"""Extra measurements next to bench.py (not the contract line): config C3 (E-step / GP fits at M = 1e5) and
config C5 (pred_magmaclust-style prediction: 10,000 new individuals onto a 2,000-point grid).

    python tools/bench_extra.py c5 [B] [T]
    python tools/bench_extra.py c3 [M]
    python tools/bench_extra.py newindiv [B] [T]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magmaclust_b200 import synth  # noqa: E402
from magmaclust_b200.engine import Engine  # noqa: E402


def c5(B=10000, T=2000, K=5, nstar=20, Tp=None):
    """trained state on a T-point grid (K = 5, M = 1000 training individuals with 30 points each), then B new
    individuals with 20 observations each, predicted on the whole grid"""
    ds = synth.make_dataset(M=1000, K=K, n=30, n_grid=T, common_hp_i=True, seed=1005)
    eng = Engine(0)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
    M = ds["M"]
    theta_k = np.tile([1.0, 1.0, 0.2], (K, 1))
    theta_i = np.tile([1.0, 1.0, 0.2], (M, 1))
    eng.set_state(theta_k, theta_i, ds["tau0"].mean(0), np.zeros(K), ds["tau0"])
    t0 = time.perf_counter()
    eng.posterior_mu_k(ds["grid_full"], tau=ds["tau0"], want_mean=False, want_cov=False)
    t_post = time.perf_counter() - t0
    rng = np.random.default_rng(7)
    obs = np.sort(np.argsort(rng.random((B, T)), axis=1)[:, :nstar], axis=1).astype(np.int32)
    mu = ds["mu_true"][rng.integers(0, K, B)]
    y = np.take_along_axis(mu, obs, axis=1) + 0.3 * rng.standard_normal((B, nstar))
    offsets = (np.arange(B + 1) * nstar).astype(np.int32)
    pred_idx = np.arange(T if Tp is None else Tp, dtype=np.int32)
    theta_new = np.array([1.0, 1.0, 0.2])
    # warm-up on a slice, then the timed call (host buffers in and out: 2 * B*K*T doubles come back)
    eng.predict(offsets[:101], obs[:100].ravel(), y[:100].ravel(), theta_new, pred_idx)
    t0 = time.perf_counter()
    tau, mean, var = eng.predict(offsets, obs.ravel(), y.ravel(), theta_new, pred_idx)
    t_e2e = time.perf_counter() - t0
    ph, nl = eng.last_timings()
    out = {"config": f"C5: {B} new individuals x {nstar} obs onto a {T}-point grid, K={K}, {len(pred_idx)} targets",
           "posterior_mu_k_s": t_post, "predict_kernel_ms": ph["predict"], "predict_e2e_s": t_e2e,
           "individuals_per_s_kernel": B / (ph["predict"] / 1e3), "individuals_per_s_e2e": B / t_e2e,
           "d2h_bytes": int(mean.nbytes + var.nbytes + tau.nbytes),
           "tau_rowsum_err": float(np.abs(tau.sum(1) - 1).max()), "var_min": float(var.min())}
    print(json.dumps(out))
    eng.close()


def c3(M=100000, K=5):
    ds = synth.make_dataset(M=M, K=K, n=30, n_grid=201, common_hp_i=True, seed=1003)
    eng = Engine(0)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
    theta_k = np.tile([1.0, 1.0, 0.2], (K, 1))
    theta_i = np.tile([1.0, 1.0, 0.2], (M, 1))
    res = {}
    for rep in range(3):
        eng.set_state(theta_k, theta_i, ds["tau0"].mean(0), np.zeros(K), ds["tau0"])
        t0 = time.perf_counter()
        eng.e_step()
        tau = eng.get_tau_new()
        res["e_step_s"] = time.perf_counter() - t0
        ph, _ = eng.last_timings()
        res["e_step_phases_ms"] = ph
    t0 = time.perf_counter()
    f, g = eng.eval_indiv_common(np.array([1.0, 1.0, 0.2]))
    res["eval_common_s"] = time.perf_counter() - t0
    ph, _ = eng.last_timings()
    res["eval_common_kernel_ms"] = ph["mstep_indiv"]
    res["gp_fits_per_s"] = M / (ph["mstep_indiv"] / 1e3)
    n = 30
    res["gp_fit_tflops"] = M * (5 * n ** 3 + 10 * n ** 2) / (ph["mstep_indiv"] / 1e3) / 1e12
    t0 = time.perf_counter()
    out = eng.vem_iteration(True, True)
    res["vem_iteration_s"] = time.perf_counter() - t0
    res["vem_phases_ms"], _ = eng.last_timings()
    res["mstep_stats"] = eng.last_mstep_stats()
    res["config"] = f"C3: M={M}, K={K}, n_i=30, N=201, common hp"
    print(json.dumps(res))
    eng.close()




def c4(N=8192, K=8, M=4096, n=32):
    """config C4: dense union grid of N timestamps, K clusters: per-cluster hyper-posterior via the multi-launch
    blocked DMMA sweep (N > 768). Checks A_k C_k = I on a few random columns against the inputs."""
    ds = synth.make_dataset(M=M, K=K, n=n, n_grid=N, common_hp_i=True, seed=1004, equispaced_cover=True)
    eng = Engine(0)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
    theta_k = np.tile([1.0, 1.0, 0.2], (K, 1))
    theta_i = np.tile([1.0, 1.0, 0.2], (M, 1))
    eng.set_state(theta_k, theta_i, ds["tau0"].mean(0), np.zeros(K), ds["tau0"])
    res = {"config": f"C4: N={N}, K={K}, M={M}, n_i={n}"}
    for rep in range(2):
        t0 = time.perf_counter()
        eng.e_step()
        res["e_step_s"] = time.perf_counter() - t0
    res["phases_ms"], res["launches"] = eng.last_timings()
    # hyper-posterior flops: N^3 per inverse, 1 distinct K_k + K A_k
    res["dense_tflops"] = (K + 1) * float(N) ** 3 / (res["phases_ms"]["dense_hyperpost"] / 1e3) / 1e12
    mean = eng.get_mean()
    tau = eng.get_tau_new()
    res["mean_finite"] = bool(np.isfinite(mean).all())
    res["tau_rowsum_err"] = float(np.abs(tau.sum(1) - 1).max())
    # residual check of one covariance: C_0 symmetric, and K^-1 consistency via a random probe of A_0 C_0 = I is not
    # available on the host without A_0; instead check symmetry and positivity of the diagonal
    c0 = eng.get_cov(0)
    res["cov_sym_err"] = float(np.abs(c0 - c0.T).max() / np.abs(c0).max())
    res["cov_diag_min"] = float(np.diag(c0).min())
    print(json.dumps(res))
    eng.close()


def newindiv(B=1000, T=2000, K=5, nstar=20):
    """free-hp EM of train_new_gp_EM (mcb_train_new_indiv) for B new individuals in one batched call, C5-shaped inputs:
    wall time, EM passes, and the log-density kernel alone (7 B evaluation points = one optimiser round)"""
    ds = synth.make_dataset(M=1000, K=K, n=30, n_grid=T, common_hp_i=True, seed=1005)
    eng = Engine(0)
    eng.set_data(ds["offsets"], ds["grid_idx_full"], ds["y"], ds["grid_full"])
    M = ds["M"]
    eng.set_state(np.tile([1.0, 1.0, 0.2], (K, 1)), np.tile([1.0, 1.0, 0.2], (M, 1)), ds["tau0"].mean(0), np.zeros(K), ds["tau0"])
    eng.posterior_mu_k(ds["grid_full"], tau=ds["tau0"], want_mean=False, want_cov=False)
    rng = np.random.default_rng(7)
    obs = np.sort(np.argsort(rng.random((B, T)), axis=1)[:, :nstar], axis=1).astype(np.int32)
    mu = ds["mu_true"][rng.integers(0, K, B)]
    y = np.take_along_axis(mu, obs, axis=1) + 0.3 * rng.standard_normal((B, nstar))
    offsets = (np.arange(B + 1) * nstar).astype(np.int32)
    theta0 = np.array([1.0, 1.0, 0.2])
    pts = np.repeat(np.arange(B, dtype=np.int32), 7)
    eng.new_indiv_logl(offsets, obs.ravel(), y.ravel(), pts[:70], np.tile(theta0, (70, 1)))          # warm-up
    t0 = time.perf_counter()
    eng.new_indiv_logl(offsets, obs.ravel(), y.ravel(), pts, np.tile(theta0, (7 * B, 1)))
    t_round = time.perf_counter() - t0
    t0 = time.perf_counter()
    th, tau, loops = eng.train_new_indiv(offsets, obs.ravel(), y.ravel(), theta0)
    t_em = time.perf_counter() - t0
    print(json.dumps({"config": f"free-hp EM: {B} new individuals x {nstar} obs, {T}-point grid, K={K}",
                      "one_round_7B_points_s": t_round, "em_s": t_em, "individuals_per_s": B / t_em,
                      "loops_mean": float(loops.mean()), "loops_max": int(loops.max()), "failed": int((loops < 0).sum()),
                      "tau_rowsum_err": float(np.nanmax(np.abs(tau.sum(1) - 1)))}))
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1]
    args = [int(a) for a in sys.argv[2:]]
    {"c5": c5, "c3": c3, "c4": c4, "newindiv": newindiv}[which](*args)
