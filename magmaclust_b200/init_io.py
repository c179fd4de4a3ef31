"""Callers and data formats either side of the hot path: the k-means initialisation of the
responsibilities (`ini_kmeans`, `ini_tau_i_k`, MAGMAclust.R:562-603) and the `db` table formats the study scripts use
(CSV with ID, Timestamp, Output columns, Simulation_study.R:554-582). The feature rows and the random choice of the
starts are host work; the clustering itself (`nstart` Lloyd runs side by side) runs on the GPU (`mcb_kmeans`,
csrc/kmeans.cu) -- there is no host k-means in the product.

Parity note: the reference clusters with `stats::kmeans(centers = k, nstart = 50)` (Hartigan-Wong, R's
RNG): there is no artefact that pins its partitions, so parity is at the level of the contract, which the tests check:
the same feature matrix (wide pivot on a common grid, otherwise (min, mean, max) per individual, MAGMAclust.R:569-580),
a local optimum of the within-cluster sum of squares, labels `K1..Kk`, and responsibilities as one-hot columns in
first-appearance order of the labels (`pivot_wider`, MAGMAclust.R:599-603, quirk Q3).
"""
from __future__ import annotations

import csv

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
def _unique_in_order(a):
    _, first = np.unique(a, return_index=True)
    return [x.item() if isinstance(x, np.generic) else x for x in np.asarray(a)[np.sort(first)]]


def kmeans_features(db):
    """The matrix `ini_kmeans` hands to kmeans (MAGMAclust.R:569-585): one row per individual in `unique(db$ID)` order.
    If every individual is observed on the same timestamps (the test `identical(unique(db$Timestamp), timestamps of the
    first ID)`) the rows are the outputs themselves (pivot_wider); otherwise (min, mean, max) of each individual."""
    ids = np.asarray(db["ID"])
    ts = np.asarray(db["Timestamp"], dtype=np.float64)
    ys = np.asarray(db["Output"], dtype=np.float64)
    uid = _unique_in_order(ids)
    rows = [np.nonzero(ids == i)[0] for i in uid]
    uts = np.asarray(_unique_in_order(ts))
    first = ts[rows[0]]
    regular = uts.shape == first.shape and np.array_equal(uts, first)
    if regular and all(len(r) == len(first) and np.array_equal(ts[r], first) for r in rows):
        feat = np.stack([ys[r] for r in rows])
    else:
        feat = np.stack([[ys[r].min(), ys[r].mean(), ys[r].max()] for r in rows])
    return uid, feat


def kmeans_starts(M, k, nstart, seed):
    """start_rows [nstart][k]: k distinct individuals per start as initial centres, like kmeans(centers = k); `seed`
    replaces R's global RNG state"""
    rng = np.random.default_rng(seed)
    return np.array([rng.choice(M, size=k, replace=False) for _ in range(max(1, int(nstart)))], dtype=np.int32)


def ini_kmeans(db, k, nstart=50, summary=False, seed=0, n_iter=100, session=None):
    """`ini_kmeans(db, k, nstart, summary)` (MAGMAclust.R:562-598): returns {"ID": [...], "Cluster_ini": ["K2", ...]}.
    `nstart` random starts (k distinct rows as centres), Lloyd iterations to a fixed point on the GPU (all starts in the
    same launches), the start with the smallest total within-cluster sum of squares wins."""
    from . import magma
    uid, feat = kmeans_features(db)
    M = feat.shape[0]
    if not (1 <= k <= M):
        raise ValueError("ini_kmeans: need 1 <= k <= number of individuals")
    s = session or magma.default_session()
    lab, cent, cost, _ = s.eng.kmeans(feat, k, kmeans_starts(M, k, nstart, seed), n_iter)
    if summary:
        print(f"K-means clustering with {k} clusters of sizes {', '.join(str(int((lab == j).sum())) for j in range(k))}; "
              f"total within-cluster sum of squares {cost:.6g}")
    return {"ID": uid, "Cluster_ini": [f"K{int(j) + 1}" for j in lab], "centers": cent, "tot_withinss": cost}


def ini_tau_i_k(db, k, nstart=50, seed=0, session=None):
    """`ini_tau_i_k(db, k, nstart)` (MAGMAclust.R:600-607): {cluster name: {ID: 0/1}}; the clusters appear in the order
    in which their labels first occur over the individuals (pivot_wider), which is NOT K1, K2, ... in general: the first
    E-step multiplies pi_k positionally against this order (quirk Q3, CF.R:757)."""
    res = ini_kmeans(db, k, nstart, seed=seed, session=session)
    names = _unique_in_order(np.asarray(res["Cluster_ini"]))
    return {name: {i: (1.0 if c == name else 0.0) for i, c in zip(res["ID"], res["Cluster_ini"])} for name in names}


# ---------------------------------------------------------------------------------------------------------------
def read_db_csv(path, dataset=None, drop_ids=()):
    """Reads the `db` columns (ID, Timestamp, Output) of a study CSV (e.g. Simulations/Data/db_i_clust_*.csv, which also
    carries Cluster, a, b, sigma, ID_dataset). `dataset` selects one `ID_dataset` value; `drop_ids` removes individuals
    (the study keeps ID 1 for prediction, Simulation_study.R:556). Extra columns are returned under their own names."""
    with open(path, newline="") as f:
        rows = list(csv.DictReader(f))
    if dataset is not None:
        rows = [r for r in rows if r.get("ID_dataset") is not None and int(float(r["ID_dataset"])) == int(dataset)]
    drop = {str(d) for d in drop_ids}
    rows = [r for r in rows if str(r["ID"]) not in drop]
    if not rows:
        raise ValueError(f"read_db_csv: no rows selected from {path}")

    def col(name, conv):
        return [conv(r[name]) for r in rows]

    def as_id(x):
        try:
            return int(x)
        except ValueError:
            return x

    db = {"ID": np.asarray(col("ID", as_id)), "Timestamp": np.asarray(col("Timestamp", float)),
          "Output": np.asarray(col("Output", float))}
    for extra in rows[0].keys():
        if extra not in db and extra != "":
            vals = col(extra, lambda x: x)
            try:
                db[extra] = np.asarray([float(v) for v in vals])
            except ValueError:
                db[extra] = np.asarray(vals)
    return db


def write_db_csv(path, db):
    """Writes ID, Timestamp, Output (+ any further equally long columns of `db`), doubles with 17 significant digits so
    that a round trip is exact."""
    cols = ["ID", "Timestamp", "Output"] + [c for c in db if c not in ("ID", "Timestamp", "Output")]
    n = len(db["ID"])
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        for r in range(n):
            w.writerow([repr(float(db[c][r])) if isinstance(db[c][r], (float, np.floating)) else db[c][r] for c in cols])
