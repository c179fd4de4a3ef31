"""Shared test helpers: golden fixture access, small synthetic problems, tolerance checks."""
import os

import numpy as np

from oracle import magma_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hoo_diff_time.npz")
NAMES3 = ["K1", "K2", "K3"]

# tolerances stated by BASELINE.json:north_star
RTOL_POST = 1e-8   # hyper-posterior means and covariances
RTOL_LL = 1e-6     # log-likelihoods and responsibilities


def load_gold():
    return np.load(GOLD)


def gold_dataset(g, d):
    """(db, hp, tau, m_k) of fixture dataset d (0-based): stored training_VEM outputs of the reference"""
    ids = g["ID"][d].astype(int)
    db = {"ID": ids, "Timestamp": g["Timestamp"][d], "Output": g["Output"][d]}
    uid = list(O.split_db(db).keys())
    hp = {"theta_k": {k: g["theta_k"][d][c] for c, k in enumerate(NAMES3)},
          "theta_i": {i: g["theta_i"][d][r] for r, i in enumerate(uid)},
          "pi_k": g["pi_k"][d]}
    tau = {k: {i: float(g["tau_i_k"][d][r, c]) for r, i in enumerate(uid)} for c, k in enumerate(NAMES3)}
    m_k = {k: 0.0 for k in NAMES3}
    return db, hp, tau, m_k


def small_problem(M=12, K=3, n=(6, 14), n_grid=40, seed=0, common_hp_i=False, ragged=True):
    """a small random problem with ragged n_i for oracle-vs-GPU unit parity"""
    rng = np.random.default_rng(seed)
    G = np.round(np.linspace(0, 10, n_grid), 3)
    ids, ts, ys = [], [], []
    centers = rng.uniform(-3, 3, K)
    for i in range(M):
        ni = int(rng.integers(n[0], n[1] + 1)) if ragged else n[1]
        t = np.sort(rng.choice(G, ni, replace=False))
        k = rng.integers(K)
        y = centers[k] + np.sin(t + k) + 0.3 * rng.standard_normal(ni)
        ids += [i + 1] * ni
        ts += list(t)
        ys += list(y)
    db = {"ID": np.array(ids), "Timestamp": np.array(ts), "Output": np.array(ys)}
    names = [f"K{k + 1}" for k in range(K)]
    uid = list(O.split_db(db).keys())
    if common_hp_i:
        th = np.tile(np.array([0.7, 0.4, 0.35]), (M, 1))
    else:
        th = np.stack([rng.uniform(0, 2, M), rng.uniform(0, 1.5, M), rng.uniform(0.2, 0.6, M)], axis=1)
    hp = {"theta_k": {k: np.array([1.2, 0.8, 0.3]) + (0.1 * c if not common_hp_i else 0.0) for c, k in enumerate(names)},
          "theta_i": {i: th[r] for r, i in enumerate(uid)}}
    w = rng.dirichlet(np.ones(K), M)
    tau = {k: {i: float(w[r, c]) for r, i in enumerate(uid)} for c, k in enumerate(names)}
    hp["pi_k"] = w.mean(0)
    m_k = {k: float(c) * 0.5 for c, k in enumerate(names)}
    return db, hp, tau, m_k


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def assert_param_close(got, ref, names_k):
    for k in names_k:
        assert rel_err(got["mean"][k]["Output"], ref["mean"][k]["Output"]) < RTOL_POST, f"mean {k}"
        assert rel_err(got["cov"][k].m, ref["cov"][k].m) < RTOL_POST, f"cov {k}"
    A, B = O.tau_to_array(got["tau_i_k"]), O.tau_to_array(ref["tau_i_k"])
    assert np.max(np.abs(A - B)) < RTOL_LL, "tau"
    assert np.array_equal(A.argmax(1), B.argmax(1)), "hard assignments"


GOLD_PRED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hoo_diff_time_pred.npz")


def gold_prediction_case(g, gp, d):
    """the reference's own prediction experiment on fixture dataset d (loop_pred, Simulation_study.R:672-737): the stored
    trained model, the testing individual's first nb_obs timestamps as observations, its last nb_test as targets"""
    db, hp, tau, m_k = gold_dataset(g, d)
    order = np.argsort(gp["Timestamp"][d], kind="stable")
    t, y = gp["Timestamp"][d][order], gp["Output"][d][order]
    nb_obs, nb_test = int(gp["nb_obs"][0]), int(gp["nb_test"][0])
    new_db = {"Timestamp": t[:nb_obs], "Output": y[:nb_obs]}
    test = {"Timestamp": t[-nb_test:], "Output": y[-nb_test:]}
    list_hp = {"hp": hp, "param": {"tau_i_k": tau}}
    return db, m_k, list_hp, new_db, test


def mse_clust(test, pred):
    """MSE_clust (Simulation_study.R:262-276): squared error of the tau-weighted mixture of the cluster means"""
    mix = sum(pred[k]["tau_k"] * np.asarray(pred[k]["Mean"]) for k in pred)
    return float(np.mean((np.asarray(test["Output"]) - mix) ** 2))


def wcic(test, pred, level=0.05):
    """WCIC (Simulation_study.R:283-298): tau-weighted share of test values inside each cluster's credible interval"""
    from scipy.stats import norm
    q = norm.ppf(1 - level / 2)
    v = np.asarray(test["Output"])
    out = 0.0
    for k in pred:
        m, sd = np.asarray(pred[k]["Mean"]), np.sqrt(np.asarray(pred[k]["Var"]))
        out += 100.0 * pred[k]["tau_k"] * np.mean((m - q * sd < v) & (v < m + q * sd))
    return float(out)
