"""Synthetic inputs for the BASELINE.json configurations (host-side data generation only).

The recipe follows the reference's own simulator, `simu_scheme` / `draw` / `prior_mean`
(Simulation_study.R:155-218): a grid G on [0, 10]; cluster mean  mu_k(t) = a_k t + b_k + f_k(t)  with
a_k ~ U[-2, 2], b_k ~ U[20, 30], f_k ~ GP(0, SE(theta1 ~ U[0,5], theta2 ~ U[0,2])); membership uniform over K;
t_i = n_i grid points sampled without replacement, sorted; y_i ~ N(mu_k(t_i), Psi_i) with
theta_i1 ~ U[0,5], theta_i2 ~ U[0,2], sigma_i ~ U[0,0.1] (shared by all individuals when `common_hp_i`);
every draw of a hyper-parameter is rounded to 2 decimals like `draw` (SS.R:155-158). R's RNG is not
available, so numpy's `default_rng(seed)` is used; the seeds are listed in BASELINE.md.

The initial responsibilities tau_0 are one-hot labels from a seeded Lloyd k-means on (min, mean, max) of each
curve, as `ini_kmeans` does for irregular timestamps (MAGMAclust.R:569-580).
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: dict(M, K, n, n_grid, common_hp_i, seed)
    "C1": dict(M=50, K=3, n=30, n_grid=201, common_hp_i=True, seed=4242),
    "C2": dict(M=1000, K=5, n=50, n_grid=401, common_hp_i=False, seed=1002),
    "C3": dict(M=100_000, K=5, n=30, n_grid=201, common_hp_i=True, seed=1003),
    "C4": dict(M=4096, K=8, n=32, n_grid=8192, common_hp_i=True, seed=1004),
    "C5": dict(M=1000, K=5, n=30, n_grid=2000, common_hp_i=True, seed=1005),
}


def _se(t, a, b):
    d = t[:, None] - t[None, :]
    return np.exp(a - 0.5 * np.exp(-b) * d * d)


def _draw(rng, lo, hi, size=None):
    return np.round(rng.uniform(lo, hi, size), 2)


def _gp_draw(rng, grid, a, b):
    """exact Cholesky draw on small grids, random Fourier features above 2048 points"""
    n = grid.shape[0]
    if n <= 2048:
        Kmat = _se(grid, a, b) + 1e-8 * np.exp(a) * np.eye(n)
        return np.linalg.cholesky(Kmat) @ rng.standard_normal(n)
    nf = 2048
    ell2 = np.exp(b)
    w = rng.standard_normal(nf) / np.sqrt(ell2)
    ph = rng.uniform(0, 2 * np.pi, nf)
    return np.sqrt(2.0 * np.exp(a) / nf) * np.cos(grid[:, None] * w[None, :] + ph[None, :]) @ rng.standard_normal(nf)


def kmeans_labels(feat, K, rng, n_start=5, n_iter=50):
    """seeded Lloyd k-means (k-means++ starts); returns labels in [0, K)"""
    best, best_cost = None, np.inf
    M = feat.shape[0]
    for _ in range(n_start):
        cent = [feat[rng.integers(M)]]
        for _k in range(1, K):
            d2 = np.min(((feat[:, None, :] - np.array(cent)[None, :, :]) ** 2).sum(-1), axis=1)
            cent.append(feat[rng.choice(M, p=d2 / d2.sum())] if d2.sum() > 0 else feat[rng.integers(M)])
        cent = np.array(cent)
        for _it in range(n_iter):
            lab = np.argmin(((feat[:, None, :] - cent[None, :, :]) ** 2).sum(-1), axis=1)
            new = np.array([feat[lab == k].mean(0) if np.any(lab == k) else cent[k] for k in range(K)])
            if np.allclose(new, cent):
                break
            cent = new
        cost = ((feat - cent[lab]) ** 2).sum()
        if cost < best_cost:
            best, best_cost = lab, cost
    return best


def make_dataset(M, K, n, n_grid, common_hp_i=True, seed=0, id_offset=0, equispaced_cover=False):
    """Returns a dict with the CSR arrays the C ABI takes plus the R-style `db` columns.

    keys: grid[N_full], offsets[M+1], grid_idx[P] (into the UNION grid of the sampled timestamps),
    union_grid[N], t[P], y[P], ids[M], true_cluster[M], tau0[M][K], theta_true_i[M][3]
    """
    rng = np.random.default_rng(seed)
    G = np.linspace(0.0, 10.0, n_grid)
    k_a, k_b = _draw(rng, 0, 5), _draw(rng, 0, 2)
    mu = np.empty((K, n_grid))
    for k in range(K):
        mu[k] = _draw(rng, -2, 2) * G + _draw(rng, 20, 30) + _gp_draw(rng, G, k_a, k_b)
    if common_hp_i:
        th = np.tile(np.array([_draw(rng, 0, 5), _draw(rng, 0, 2), _draw(rng, 0, 0.1)]), (M, 1))
    else:
        th = np.stack([_draw(rng, 0, 5, M), _draw(rng, 0, 2, M), _draw(rng, 0, 0.1, M)], axis=1)
    clus = rng.integers(0, K, M)
    # n distinct sorted grid points per individual
    if equispaced_cover:
        # C4: M*n >= N; a random permutation of the grid dealt out n points at a time touches every point
        perm = rng.permutation(n_grid)
        gi = np.sort(perm[(np.arange(M)[:, None] * n + np.arange(n)[None, :]) % n_grid], axis=1)
    else:
        gi = np.empty((M, n), dtype=np.int64)
        chunk = 4096
        for s in range(0, M, chunk):
            e = min(M, s + chunk)
            gi[s:e] = np.sort(np.argsort(rng.random((e - s, n_grid)), axis=1)[:, :n], axis=1)
    t = G[gi]
    y = np.empty((M, n))
    chunk = 4096
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        d = t[s:e, :, None] - t[s:e, None, :]
        Psi = np.exp(th[s:e, 0, None, None] - 0.5 * np.exp(-th[s:e, 1, None, None]) * d * d)
        Psi += (th[s:e, 2] ** 2 + 1e-10)[:, None, None] * np.eye(n)[None]
        L = np.linalg.cholesky(Psi)
        z = rng.standard_normal((e - s, n, 1))
        y[s:e] = np.take_along_axis(mu[clus[s:e]], gi[s:e], axis=1) + (L @ z)[..., 0]
    used = np.unique(gi)
    remap = -np.ones(n_grid, dtype=np.int64)
    remap[used] = np.arange(used.shape[0])
    feat = np.stack([y.min(1), y.mean(1), y.max(1)], axis=1)
    lab = kmeans_labels(feat, K, rng) if M <= 20000 else _kmeans_big(feat, K, rng)
    tau0 = np.zeros((M, K))
    tau0[np.arange(M), lab] = 1.0
    ids = np.arange(M) + 1 + id_offset
    return {
        "grid_full": G,
        "union_grid": G[used],
        "offsets": (np.arange(M + 1) * n).astype(np.int32),
        "grid_idx": remap[gi].ravel().astype(np.int32),
        "grid_idx_full": gi.ravel().astype(np.int32),
        "t": t.ravel(),
        "y": y.ravel(),
        "ids": ids,
        "true_cluster": clus,
        "tau0": tau0,
        "theta_true_i": th,
        "mu_true": mu,
        "M": M, "K": K, "n": n, "N": int(used.shape[0]),
    }


def _kmeans_big(feat, K, rng):
    """k-means on a subsample, labels for everyone by nearest centre (keeps generation fast at M = 1e5+)"""
    sub = rng.choice(feat.shape[0], 20000, replace=False)
    lab = kmeans_labels(feat[sub], K, rng, n_start=3)
    cent = np.array([feat[sub][lab == k].mean(0) for k in range(K)])
    out = np.empty(feat.shape[0], dtype=np.int64)
    for s in range(0, feat.shape[0], 65536):
        e = min(feat.shape[0], s + 65536)
        out[s:e] = np.argmin(((feat[s:e, None, :] - cent[None]) ** 2).sum(-1), axis=1)
    return out


def shard(ds, rank, nranks):
    """contiguous block of individuals for one rank (indices into the FULL grid so that every rank shares it)"""
    M = ds["M"]
    per = (M + nranks - 1) // nranks
    lo, hi = min(M, rank * per), min(M, (rank + 1) * per)
    o = ds["offsets"]
    a, b = int(o[lo]), int(o[hi])
    return {"grid": ds["grid_full"], "offsets": (o[lo:hi + 1] - o[lo]).astype(np.int32),
            "grid_idx": ds["grid_idx_full"][a:b].copy(), "y": ds["y"][a:b].copy(), "t": ds["t"][a:b].copy(),
            "ids": ds["ids"][lo:hi], "tau0": ds["tau0"][lo:hi].copy(), "M": hi - lo, "M_total": M,
            "K": ds["K"], "N": int(ds["grid_full"].shape[0]), "n": ds["n"]}


def to_db(ds):
    """R-style tibble columns for the oracle / the reference-shaped API"""
    n = np.diff(ds["offsets"])
    return {"ID": np.repeat(ds["ids"], n), "Timestamp": ds["t"].copy(), "Output": ds["y"].copy()}


def tau_dict(ds, names_k=None):
    K = ds["K"]
    names_k = names_k or [f"K{k + 1}" for k in range(K)]
    return {k: {int(i): float(ds["tau0"][r, c]) for r, i in enumerate(ds["ids"])} for c, k in enumerate(names_k)}


def config(name, M=None, seed=None):
    c = dict(CONFIGS[name])
    if M is not None:
        c["M"] = M
    if seed is not None:
        c["seed"] = seed
    return make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"],
                        equispaced_cover=(name == "C4"))
