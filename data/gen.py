"""Synthetic inputs of the BASELINE.json configurations, written ONCE to files so that the untouched R reference
(baseline/run_reference.R), the CPU oracle and the GPU path read identical bytes.

    python data/gen.py C1 [C2 ...] [--out data/generated] [--individuals M]

Per configuration directory:
    db.csv      ID,Timestamp,Output,Cluster      the reference's training-set schema (Simulations/Data/*.csv:1); numbers are
                                                 printed with repr() (shortest round-trip form): read back they are the same doubles
    tau0.csv    ID,K1..KK                        initial responsibilities (one-hot labels of the seeded k-means on (min, mean, max),
                                                 MAGMAclust.R:569-607); R: ini_tau_i_k = lapply(tau0[-1], function(c) setNames(as.list(c), tau0$ID))
    csr.bin     int32 M, N, K, P, n_full | int32 offsets[M+1] | int32 grid_idx[P] (into the union grid) |
                int32 grid_idx_full[P] (into the full grid) | f64 t[P] | f64 y[P] | f64 union_grid[N] | f64 grid_full[n_full] |
                f64 tau0[M][K] | int32 ids[M] | int32 true_cluster[M]        (little endian; what mcb_set_data takes)
    meta.json   configuration, seed, shapes, sha256 of the three files

The generator itself is magmaclust_b200/synth.py (recipe of simu_scheme, Simulation_study.R:155-218); `load()` returns the same
dict synth.make_dataset returns, so bench.py --data DIR and the tests consume the files instead of regenerating.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def write(ds, cfg, out_dir):
    """write one dataset (a dict of synth.make_dataset) as db.csv, tau0.csv, csr.bin, meta.json"""
    os.makedirs(out_dir, exist_ok=True)
    M, K, N = int(ds["M"]), int(ds["K"]), int(ds["N"])
    n = np.diff(ds["offsets"])
    ids = np.repeat(ds["ids"], n)
    clus = np.repeat(ds["true_cluster"], n)
    with open(os.path.join(out_dir, "db.csv"), "w") as f:
        f.write("ID,Timestamp,Output,Cluster\n")
        f.writelines(f"{int(i)},{float(t)!r},{float(y)!r},K{int(c) + 1}\n" for i, t, y, c in zip(ids, ds["t"], ds["y"], clus))
    with open(os.path.join(out_dir, "tau0.csv"), "w") as f:
        f.write("ID," + ",".join(f"K{k + 1}" for k in range(K)) + "\n")
        f.writelines(f"{int(i)}," + ",".join(repr(float(v)) for v in row) + "\n" for i, row in zip(ds["ids"], ds["tau0"]))
    P = int(ds["y"].shape[0])
    with open(os.path.join(out_dir, "csr.bin"), "wb") as f:
        f.write(np.array([M, N, K, P, ds["grid_full"].shape[0]], dtype="<i4").tobytes())
        for a, dt in ((ds["offsets"], "<i4"), (ds["grid_idx"], "<i4"), (ds["grid_idx_full"], "<i4"), (ds["t"], "<f8"),
                      (ds["y"], "<f8"), (ds["union_grid"], "<f8"), (ds["grid_full"], "<f8"), (ds["tau0"], "<f8"),
                      (ds["ids"], "<i4"), (ds["true_cluster"], "<i4")):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
    meta = {"config": cfg, "M": M, "K": K, "N": N, "P": P, "n": int(ds["n"]),
            "start_state": {"theta_k": [1.0, 1.0, 0.2], "theta_i": [1.0, 1.0, 0.2], "m_k": 0.0},
            "sha256": {fn: _sha(os.path.join(out_dir, fn)) for fn in ("db.csv", "tau0.csv", "csr.bin")}}
    with open(os.path.join(out_dir, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


def load(out_dir, verify=True):
    """the dict synth.make_dataset returns, from csr.bin (hashes of meta.json checked unless verify = False)"""
    meta = json.load(open(os.path.join(out_dir, "meta.json")))
    path = os.path.join(out_dir, "csr.bin")
    if verify and _sha(path) != meta["sha256"]["csr.bin"]:
        raise ValueError(f"{path}: content does not match meta.json")
    raw = open(path, "rb").read()
    M, N, K, P, nf = (int(v) for v in np.frombuffer(raw, "<i4", 5))
    o = 20

    def take(count, dt):
        nonlocal o
        a = np.frombuffer(raw, dt, count, o).copy()
        o += a.nbytes
        return a
    ds = {"offsets": take(M + 1, "<i4"), "grid_idx": take(P, "<i4"), "grid_idx_full": take(P, "<i4"), "t": take(P, "<f8"),
          "y": take(P, "<f8"), "union_grid": take(N, "<f8"), "grid_full": take(nf, "<f8"), "tau0": take(M * K, "<f8").reshape(M, K),
          "ids": take(M, "<i4").astype(np.int64), "true_cluster": take(M, "<i4").astype(np.int64)}
    assert o == len(raw)
    ds.update(M=M, K=K, N=N, n=meta["n"])
    return ds, meta["config"]


def load_db_csv(out_dir):
    """db.csv as the R-style column dict (what the reference reads with read_csv): proves the CSV carries the same doubles"""
    ids, ts, ys = [], [], []
    with open(os.path.join(out_dir, "db.csv")) as f:
        next(f)
        for ln in f:
            a, b, c, _ = ln.split(",")
            ids.append(int(a)); ts.append(float(b)); ys.append(float(c))
    return {"ID": np.array(ids), "Timestamp": np.array(ts), "Output": np.array(ys)}


def main():
    from magmaclust_b200 import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=sorted(synth.CONFIGS))
    ap.add_argument("--out", default=os.path.join(ROOT, "data", "generated"))
    ap.add_argument("--individuals", type=int, default=0, help="override M (e.g. a reduced C3 for the R reference)")
    a = ap.parse_args()
    for name in a.configs:
        c = dict(synth.CONFIGS[name])
        if a.individuals:
            c["M"] = a.individuals
        ds = synth.make_dataset(c["M"], c["K"], c["n"], c["n_grid"], c["common_hp_i"], c["seed"], equispaced_cover=(name == "C4"))
        d = os.path.join(a.out, name if not a.individuals else f"{name}_M{a.individuals}")
        meta = write(ds, dict(c, name=name), d)
        print(d, {k: meta[k] for k in ("M", "K", "N", "P")})


if __name__ == "__main__":
    main()
