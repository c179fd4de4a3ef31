"""Regenerates tests/golden/hoo_diff_time.npz from the reference's own stored study outputs.

Run in the BUILD container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Source pair: inputs `Simulations/Data/db_i_clust_100rep_M20_N30_diff_time_Hoo.csv`
(training individuals are IDs 2..21 of each dataset, Simulation_study.R:554-582) and the outputs that
`training_VEM` itself produced for them, `Simulations/Training/train_for_pred_Hoo_diff_time.rds`
(`[[d]]$MAGMAclust`: hp, Time_train, tau_i_k). Nothing is computed here: the arrays are re-encoded verbatim.
"""
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from tools.rds import read_rds  # noqa: E402

REF = "/root/reference/Simulations"


def main():
    rds = read_rds(f"{REF}/Training/train_for_pred_Hoo_diff_time.rds")
    csv = pd.read_csv(f"{REF}/Data/db_i_clust_100rep_M20_N30_diff_time_Hoo.csv")
    csv = csv[csv.ID != 1]
    D = 100
    ids = np.zeros((D, 600), np.int16)
    ts = np.zeros((D, 600))
    ys = np.zeros((D, 600))
    true_cluster = np.zeros((D, 20), np.int8)
    theta_k = np.zeros((D, 3, 3))
    theta_i = np.zeros((D, 20, 3))
    pi_k = np.zeros((D, 3))
    tau = np.zeros((D, 20, 3))
    time_train = np.zeros(D)
    for d in range(D):
        df = csv[csv.ID_dataset == d + 1]
        assert len(df) == 600
        ids[d], ts[d], ys[d] = df.ID.values, df.Timestamp.values, df.Output.values
        mc = rds[d]["MAGMAclust"]
        hp = mc["hp"]
        names_k = hp["theta_k"].names
        names_i = hp["theta_i"].names
        assert names_k == ["K1", "K2", "K3"] and names_i == [str(i) for i in range(2, 22)]
        first = df.groupby("ID", sort=False).Cluster.first()
        true_cluster[d] = [int(first[i][1:]) for i in range(2, 22)]
        for c, k in enumerate(names_k):
            theta_k[d, c] = np.asarray(hp["theta_k"][k])
            for r, i in enumerate(names_i):
                tau[d, r, c] = float(mc["tau_i_k"][k][i][0])
        for r, i in enumerate(names_i):
            theta_i[d, r] = [col[0] for col in hp["theta_i"][i]]
        pi_k[d] = np.asarray(hp["pi_k"])
        time_train[d] = float(mc["Time_train"][0])
    settings = {n: rds[n] for n in rds.names[100:]}
    assert bool(settings["common_hp_k"][0]) and bool(settings["common_hp_i"][0])
    out = os.path.join(os.path.dirname(__file__), "hoo_diff_time.npz")
    np.savez_compressed(out, ID=ids, Timestamp=ts, Output=ys, true_cluster=true_cluster, theta_k=theta_k,
                        theta_i=theta_i, pi_k=pi_k, tau_i_k=tau, Time_train=time_train,
                        ini_hp_theta_k=np.asarray(settings["ini_hp_clust"][0]),
                        ini_hp_theta_i=np.asarray(settings["ini_hp_clust"][1]),
                        prior_mean=np.asarray(settings["prior_mean"]))
    print("wrote", out, os.path.getsize(out), "bytes")
    make_pred()


def make_pred():
    """hoo_diff_time_pred.npz: the testing individual (ID 1) of every dataset and the numbers the reference's own study
    stored for its MAGMAclust prediction (`loop_pred`, Simulation_study.R:672-737: first 20 timestamps observed, last 10
    predicted, `full_algo_clust` with the stored trained model, then MSE_clust / WCIC, Simulation_study.R:262-298):
    `Simulations/Results/res_pred_Hoo_diff_time.csv`, rows Method == "MAGMAclust" in dataset order. Re-encoded verbatim."""
    csv = pd.read_csv(f"{REF}/Data/db_i_clust_100rep_M20_N30_diff_time_Hoo.csv")
    res = pd.read_csv(f"{REF}/Results/res_pred_Hoo_diff_time.csv")
    mc = res[res.Method == "MAGMAclust"].reset_index(drop=True)
    assert len(mc) == 100
    one = csv[csv.ID == 1]
    ts, ys = np.zeros((100, 30)), np.zeros((100, 30))
    for d in range(100):
        df = one[one.ID_dataset == d + 1]
        assert len(df) == 30
        ts[d], ys[d] = df.Timestamp.values, df.Output.values
    out = os.path.join(os.path.dirname(__file__), "hoo_diff_time_pred.npz")
    np.savez_compressed(out, Timestamp=ts, Output=ys, MSE=mc.MSE.values, WCIC=mc.WCIC.values,
                        Time_train=mc.Time_train.values, nb_obs=np.array([20]), nb_test=np.array([10]))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
