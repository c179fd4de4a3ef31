"""Rows of SURVEY.md §8f either side of the path: ini_kmeans / ini_tau_i_k (MAGMAclust.R:562-607; features and starts on the
host, the clustering on the GPU, checked bit for bit against oracle/kmeans_oracle.py) and the CSV form of `db`."""
import os

import numpy as np
import pytest

from magmaclust_b200 import init_io, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "hoo_diff_time.npz")


def _db(regular, seed=3, M=24, K=3, n=10):
    ds = synth.make_dataset(M=M, K=K, n=n, n_grid=n if regular else 40, common_hp_i=True, seed=seed)
    return synth.to_db(ds), ds


def test_features_follow_the_reference_rule():
    db, _ = _db(regular=True)
    uid, feat = init_io.kmeans_features(db)
    assert feat.shape == (24, 10)                      # common grid: pivot_wider of the outputs
    ids = np.asarray(db["ID"])
    assert np.array_equal(feat[0], np.asarray(db["Output"])[ids == uid[0]])
    db2, _ = _db(regular=False)
    uid2, feat2 = init_io.kmeans_features(db2)
    assert feat2.shape == (24, 3)                      # irregular: (min, mean, max), MAGMAclust.R:571-579
    y0 = np.asarray(db2["Output"])[np.asarray(db2["ID"]) == uid2[0]]
    assert np.allclose(feat2[0], [y0.min(), y0.mean(), y0.max()])


@pytest.mark.parametrize("regular", [True, False])
def test_oracle_kmeans_is_a_lloyd_fixed_point(regular):
    """CPU: the restated k-means ends where every individual sits with its nearest centre and every centre is its cluster's mean"""
    from oracle import kmeans_oracle as KO
    db, _ = _db(regular)
    uid, feat = init_io.kmeans_features(db)
    lab, cent, cost, iters = KO.kmeans(feat, 3, init_io.kmeans_starts(feat.shape[0], 3, 20, 1))
    for k in range(3):
        if np.any(lab == k):
            assert np.allclose(cent[k], feat[lab == k].mean(0), rtol=1e-14)
    d2 = ((feat[:, None, :] - cent[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(np.argmin(d2, 1), lab) and np.isclose(cost, d2[np.arange(len(lab)), lab].sum(), rtol=1e-13)
    assert 1 <= iters <= 100


@pytest.mark.gpu
@pytest.mark.parametrize("M,D,K,nstart", [(24, 3, 3, 20), (24, 10, 3, 7), (2500, 3, 5, 12), (1, 3, 1, 2), (40, 64, 4, 5)])
def test_gpu_kmeans_is_bit_identical_to_the_oracle(M, D, K, nstart):
    """mcb_kmeans vs oracle/kmeans_oracle.py on the same features and starts: identical labels, centres and cost (every
    floating-point operation has a prescribed order), including M > one summation chunk and a single individual"""
    from magmaclust_b200.engine import Engine
    from oracle import kmeans_oracle as KO
    rng = np.random.default_rng(M + D)
    feat = rng.normal(0, 1, (M, D)) + 4.0 * rng.integers(0, K, M)[:, None]
    starts = init_io.kmeans_starts(M, K, nstart, 3)
    e = Engine(0)
    try:
        lab, cent, cost, iters = e.kmeans(feat, K, starts)
    finally:
        e.close()
    lab_o, cent_o, cost_o, iters_o = KO.kmeans(feat, K, starts)
    assert np.array_equal(lab, lab_o) and np.array_equal(cent, cent_o) and cost == cost_o and iters == iters_o


@pytest.mark.gpu
@pytest.mark.parametrize("regular", [True, False])
def test_ini_kmeans_is_a_local_optimum_and_labels_are_K_names(regular):
    db, _ = _db(regular)
    res = init_io.ini_kmeans(db, 3, nstart=20, seed=1)
    uid, feat = init_io.kmeans_features(db)
    assert res["ID"] == uid and set(res["Cluster_ini"]) <= {"K1", "K2", "K3"}
    lab = np.array([int(c[1:]) - 1 for c in res["Cluster_ini"]])
    cent = res["centers"]
    for k in range(3):
        if np.any(lab == k):
            assert np.allclose(cent[k], feat[lab == k].mean(0))
    d2 = ((feat[:, None, :] - cent[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(np.argmin(d2, 1), lab)       # every individual sits with its nearest centre (Lloyd fixed point)
    assert np.isclose(res["tot_withinss"], d2[np.arange(len(lab)), lab].sum())
    # more starts never give a worse optimum (same seed: the first 20 starts are shared)
    assert init_io.ini_kmeans(db, 3, nstart=40, seed=1)["tot_withinss"] <= res["tot_withinss"] + 1e-12


@pytest.mark.gpu
def test_ini_tau_i_k_is_one_hot_in_first_appearance_order():
    db, _ = _db(regular=False, seed=8)
    tau = init_io.ini_tau_i_k(db, 3, nstart=10, seed=2)
    res = init_io.ini_kmeans(db, 3, nstart=10, seed=2)
    seen = []
    for c in res["Cluster_ini"]:
        if c not in seen:
            seen.append(c)
    assert list(tau.keys()) == seen                    # pivot_wider column order, quirk Q3
    for i, c in zip(res["ID"], res["Cluster_ini"]):
        assert [tau[k][i] for k in tau] == [1.0 if k == c else 0.0 for k in tau]


@pytest.mark.gpu
def test_well_separated_clusters_are_recovered():
    rng = np.random.default_rng(0)
    ids, ts, ys, truth = [], [], [], []
    for i in range(30):
        k = i % 3
        t = np.sort(rng.choice(np.arange(0, 10, 0.25), 8, replace=False))
        ids += [i + 1] * 8
        ts += list(t)
        ys += list(20.0 * k + rng.normal(0, 0.3, 8))
        truth.append(k)
    res = init_io.ini_kmeans({"ID": np.array(ids), "Timestamp": np.array(ts), "Output": np.array(ys)}, 3, nstart=10)
    lab = np.array([int(c[1:]) for c in res["Cluster_ini"]])
    for k in range(3):                                  # a relabelling of the truth
        assert len(set(lab[np.array(truth) == k])) == 1
    assert len(set(lab)) == 3


def test_csv_round_trip_is_exact(tmp_path):
    db, _ = _db(regular=False, seed=5)
    p = tmp_path / "db.csv"
    init_io.write_db_csv(p, db)
    back = init_io.read_db_csv(p)
    assert np.array_equal(back["ID"], np.asarray(db["ID"]))
    assert np.array_equal(back["Timestamp"], np.asarray(db["Timestamp"]))   # bit-exact: repr() of a double round-trips
    assert np.array_equal(back["Output"], np.asarray(db["Output"]))


def test_study_csv_layout(tmp_path):
    """the layout of Simulations/Data/db_i_clust_*.csv: extra columns, several data sets, ID 1 kept for prediction"""
    p = tmp_path / "study.csv"
    with open(p, "w") as f:
        f.write("ID,Timestamp,Output,Cluster,a,b,sigma,ID_dataset\n")
        for d in (1, 2):
            for i in (1, 2, 3):
                for t in (0.35000000000000003, 0.45):
                    f.write(f"{i},{t!r},{16.5 + i + d},K{i},4.03,1.1,0.61,{d}\n")
    db = init_io.read_db_csv(p, dataset=2, drop_ids=(1,))
    assert list(db["ID"]) == [2, 2, 3, 3] and db["Timestamp"][0] == 0.35000000000000003
    assert list(db["Cluster"]) == ["K2", "K2", "K3", "K3"] and np.all(db["ID_dataset"] == 2)
