"""data/gen.py: the files the R reference (baseline/run_reference.R), the oracle and the GPU path all read."""
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _formals(text, fn):
    """formal argument names of `fn = function(...)` / `fn <- function (...)` in R source text"""
    m = re.search(r"^" + fn + r"\s*(?:=|<-)\s*function\s*\(", text, re.M)
    assert m, fn
    i, depth, cur, out = m.end(), 1, "", []
    while depth > 0:
        ch = text[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
            if depth == 0:
                break
        if ch == "," and depth == 1:
            out.append(cur); cur = ""
        else:
            cur += ch
        i += 1
    out.append(cur)
    return [a.split("=")[0].strip() for a in out if a.strip()]


def test_files_round_trip_to_the_generators_arrays(tmp_path):
    """csr.bin gives back exactly what synth.make_dataset produced; db.csv (what R's read_csv sees) holds the same doubles in
    the same row order; tau0.csv rows are one-hot; meta.json hashes match; a tampered file is refused"""
    from data import gen
    from magmaclust_b200 import synth
    out = subprocess.run([sys.executable, os.path.join(ROOT, "data", "gen.py"), "C1", "--out", str(tmp_path)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    d = str(tmp_path / "C1")
    ds, cfg = gen.load(d)
    ref = synth.config("C1")
    for k in ("offsets", "grid_idx", "grid_idx_full", "t", "y", "union_grid", "grid_full", "tau0", "ids", "true_cluster"):
        assert np.array_equal(ds[k], ref[k]), k
    assert (ds["M"], ds["K"], ds["N"], ds["n"]) == (ref["M"], ref["K"], ref["N"], ref["n"]) and cfg["name"] == "C1"
    db = gen.load_db_csv(d)
    want = synth.to_db(ref)
    assert np.array_equal(db["ID"], want["ID"]) and np.array_equal(db["Timestamp"], want["Timestamp"])
    assert np.array_equal(db["Output"], want["Output"])            # bit-identical through the decimal text
    rows = [ln.split(",") for ln in open(os.path.join(d, "tau0.csv")).read().split()[1:]]
    assert len(rows) == ref["M"] and all(sum(float(v) for v in r[1:]) == 1.0 for r in rows)
    with open(os.path.join(d, "csr.bin"), "r+b") as f:
        f.seek(100)
        b = f.read(1)
        f.seek(100)
        f.write(bytes([b[0] ^ 1]))
    try:
        gen.load(d)
        raise AssertionError("tampered csr.bin accepted")
    except ValueError:
        pass


def test_run_reference_script_calls_only_reference_functions_with_their_signatures():
    """baseline/run_reference.R cannot run here (no R): at least every function it calls exists in the reference with the
    argument count used, so the script is not written against an imagined API"""
    src = open(os.path.join(ROOT, "baseline", "run_reference.R")).read()
    refdir = "/root/reference"
    if not os.path.isdir(refdir):
        import pytest
        pytest.skip("reference sources not present on this box")
    ref = open(os.path.join(refdir, "MAGMAclust.R")).read() + open(os.path.join(refdir, "Computing_functions.R")).read()
    for fn, nargs in (("e_step_VEM", 6), ("logL_monitoring_VEM", 6), ("m_step_VEM", 8), ("training_VEM", 8),
                      ("posterior_mu_k", 6), ("train_new_gp_EM", 5), ("pred_gp_clust", 5), ("pred_max_cluster", 1)):
        assert len(_formals(ref, fn)) == nargs, fn
        assert re.search(r"\b" + fn + r"\(", src), fn


def test_r_wrappers_define_every_function_of_the_path_with_the_reference_argument_names():
    """r/R/magmaclust_b200.R: same function names and formal argument names as the reference definitions it masks"""
    refdir = "/root/reference"
    if not os.path.isdir(refdir):
        import pytest
        pytest.skip("reference sources not present on this box")
    ref = open(os.path.join(refdir, "MAGMAclust.R")).read() + open(os.path.join(refdir, "Computing_functions.R")).read()
    ours = open(os.path.join(ROOT, "r", "R", "magmaclust_b200.R")).read()

    names = ["kern_to_cov", "kern_to_inv", "dmvnorm", "logL_GP_mod", "logL_clust_multi_GP", "logL_GP_mod_common_hp_k",
             "logL_clust_multi_GP_common_hp_i", "logL_monitoring_VEM", "BIC", "gr_GP_mod", "gr_clust_multi_GP",
             "gr_GP_mod_common_hp_k", "gr_clust_multi_GP_common_hp_i", "e_step_VEM", "m_step_VEM", "update_tau_i_k_VEM",
             "update_tau_star_k_EM", "training_VEM", "model_selection", "train_new_gp_EM", "posterior_mu_k", "pred_gp_clust",
             "full_algo_clust", "pred_max_cluster"]
    for fn in names:
        assert _formals(ours, fn) == _formals(ref, fn), fn
