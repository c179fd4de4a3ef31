"""End-to-end: training_VEM (MC.R:16-99) and the prediction chain (MC.R:196-274, CF.R:764-786)."""
import numpy as np
import pytest

from oracle import magma_oracle as O
from helpers import NAMES3, RTOL_LL, RTOL_POST, gold_dataset, load_gold, rel_err, small_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,perturb", [(0, False), (2, False), (2, True)])
def test_training_golden_dataset(session, d, perturb):
    """Train from the study's start state (theta = (1,1,0.2), prior mean 0; the ini_hp_clust stored in the
    fixture) with tau_0 = one-hot of the reference's stored clustering: the GPU run follows the oracle's run
    (same iteration count, ELBO trace, hard assignments) and ends on the reference's stored clustering. With a
    perturbed start only GPU-vs-oracle agreement is required."""
    from magmaclust_b200 import magma
    g = load_gold()
    db, hp_stored, tau_stored, m_k = gold_dataset(g, d)
    ids = list(O.split_db(db).keys())
    lab = O.tau_to_array(tau_stored).argmax(1)
    lab0 = lab.copy()
    if perturb:
        lab0[0] = (lab0[0] + 1) % 3
        lab0[7] = (lab0[7] + 2) % 3
    tau0 = {k: {i: 1.0 if lab0[r] == c else 0.0 for r, i in enumerate(ids)} for c, k in enumerate(NAMES3)}
    ini = {"theta_k": g["ini_hp_theta_k"], "theta_i": g["ini_hp_theta_i"]}
    ref = O.training_VEM(db, m_k, ini, O.kernel_mu, O.kernel, tau0, True, True)
    got = magma.training_VEM(db, m_k, ini, magma.kernel_mu, magma.kernel, tau0, True, True, session=session)
    assert got["convergence"] == ref["convergence"]
    assert len(got["logL_iterations"]) == len(ref["logL_iterations"])
    hard_got = magma.pred_max_cluster(got["param"]["tau_i_k"])
    assert np.array_equal(hard_got, O.pred_max_cluster(ref["param"]["tau_i_k"]))
    if not perturb:
        assert np.array_equal(hard_got, lab + 1)  # the reference's stored clustering
    assert np.allclose(got["logL_iterations"], ref["logL_iterations"], rtol=1e-4)
    for k in NAMES3:
        assert np.allclose(got["hp"]["theta_k"][k], ref["hp"]["theta_k"][k], atol=2e-2)


def test_prediction_chain(session):
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=10, seed=9, common_hp_i=True)
    names = list(m_k.keys())
    ref_param = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    list_hp_ref = {"hp": hp, "param": ref_param}
    new_db = {"ID": np.array([99] * 5), "Timestamp": np.array([0.513, 2.051, 4.103, 6.154, 8.205]),
              "Output": np.array([0.3, 1.1, -0.2, 0.4, 0.9])}
    targets = np.round(np.linspace(0.1, 9.9, 23), 3)
    t_pred = np.unique(np.r_[targets, db["Timestamp"], new_db["Timestamp"]])
    mu_ref = O.posterior_mu_k(db, t_pred, m_k, O.kernel_mu, O.kernel, list_hp_ref)
    mu_got = magma.posterior_mu_k(db, t_pred, m_k, magma.kernel_mu, magma.kernel, list_hp_ref, session=session)
    for k in names:
        assert rel_err(mu_got["mean"][k]["Output"], mu_ref["mean"][k]["Output"]) < RTOL_POST
        assert rel_err(mu_got["cov"][k].m, mu_ref["cov"][k].m) < RTOL_POST
    hn_ref = O.train_new_gp_EM(new_db, mu_ref, None, O.kernel, hp_i=list_hp_ref)
    hn_got = magma.train_new_gp_EM(new_db, mu_got, None, magma.kernel, hp_i=list_hp_ref, session=session)
    for k in names:
        assert abs(hn_got["tau_k"][k] - hn_ref["tau_k"][k]) < RTOL_LL
    p_ref = O.pred_gp_clust(new_db, targets, mu_ref, O.kernel, hn_ref)
    p_got = magma.pred_gp_clust(new_db, targets, mu_got, magma.kernel, hn_got, session=session)
    for k in names:
        assert rel_err(p_got[k]["Mean"], p_ref[k]["Mean"]) < RTOL_POST
        assert np.max(np.abs(p_got[k]["Var"] - p_ref[k]["Var"])) < 1e-8 * max(1.0, np.max(np.abs(p_ref[k]["Var"])))
