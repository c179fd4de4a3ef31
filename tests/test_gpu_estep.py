"""GPU parity of the E-step (e_step_VEM, CF.R:589-629) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

from oracle import magma_oracle as O
from helpers import NAMES3, RTOL_LL, assert_param_close, gold_dataset, load_gold, rel_err, small_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [0, 1, 2, 29, 57])
def test_estep_golden_fixture(session, d):
    """stored reference outputs (hp, tau) of fixture dataset d as inputs: GPU E-step == oracle E-step"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = gold_dataset(load_gold(), d)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau, return_logL=True)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert_param_close(got, ref, NAMES3)
    assert rel_err(got["mat_logL"], ref["mat_logL"]) < RTOL_LL
    # the reference's stored tau is a near fixed point of the E-step: identical hard assignments
    stored = O.tau_to_array(tau)
    assert np.array_equal(O.tau_to_array(got["tau_i_k"]).argmax(1), stored.argmax(1))


@pytest.mark.parametrize("seed,common", [(0, False), (1, True), (2, False)])
def test_estep_small_ragged(session, seed, common):
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(seed=seed, common_hp_i=common)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau, return_logL=True)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert_param_close(got, ref, list(m_k.keys()))
    assert rel_err(got["mat_logL"], ref["mat_logL"]) < RTOL_LL


def test_estep_single_observation_individual(session):
    """n_i = 1 takes the scalar branch of kern_to_inv (CF.R:167-170)"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=8, n=(1, 5), seed=5)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert_param_close(got, ref, list(m_k.keys()))


def test_estep_one_hot_tau_and_vector_prior(session):
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=10, seed=7)
    ids = list(O.split_db(db).keys())
    names = list(m_k.keys())
    tau = {k: {i: 1.0 if (r % len(names)) == c else 0.0 for r, i in enumerate(ids)} for c, k in enumerate(names)}
    N = np.unique(db["Timestamp"]).shape[0]
    m_k = {k: np.linspace(0, 1, N) * (c + 1) for c, k in enumerate(names)}
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert_param_close(got, ref, names)


def test_estep_large_grid_takes_the_multi_launch_inverse(session):
    """N > 768 union timestamps: the N x N solves leave the shared-memory slab kernel for the multi-launch blocked sweep
    (128 pivots per pass over the matrix: panel sweeps + one rank-128 DMMA update, dense.cu)"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=90, K=2, n=(25, 30), n_grid=1000, seed=13)
    N = np.unique(db["Timestamp"]).shape[0]
    assert N > 768
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau, return_logL=True)
    got = magma.e_step_VEM(db, m_k, magma.kernel_mu, magma.kernel, hp, tau, session=session)
    assert_param_close(got, ref, list(m_k.keys()))
    assert rel_err(got["mat_logL"], ref["mat_logL"]) < RTOL_LL


def test_update_tau_i_k_VEM_standalone(session):
    """CF.R:733-762 as its own call: responsibilities from GIVEN means / covariances (here the oracle's E-step output at
    other hyper-parameters than the ones the responsibilities are evaluated with)"""
    from magmaclust_b200 import magma
    db, hp, tau, m_k = small_problem(M=9, seed=12)
    ref = O.e_step_VEM(db, m_k, O.kernel_mu, O.kernel, hp, tau)
    hp2 = dict(hp, theta_i={i: np.asarray(v) + 0.1 for i, v in hp["theta_i"].items()})
    pi = np.array([0.5, 0.3, 0.2])
    t_ref, _ = O.update_tau_i_k_VEM(db, m_k, ref["mean"], ref["cov"], O.kernel, hp2, pi)
    t_got = magma.update_tau_i_k_VEM(db, m_k, ref["mean"], ref["cov"], magma.kernel, hp2, pi, session=session)
    A, B = O.tau_to_array(t_got), O.tau_to_array(t_ref)
    assert np.max(np.abs(A - B)) < RTOL_LL
    assert np.array_equal(A.argmax(1), B.argmax(1))
