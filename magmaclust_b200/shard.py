"""Host-side sharding plan for the multi-GPU path (one process per GPU).

Individuals are conditionally independent given the hyper-posteriors (mean_k, cov_k); the only coupling in
e_step_VEM is the sums A_k = K_k^-1 + sum_i tau_ik P_i' Psi_i^-1 P_i and w_k = K_k^-1 m_k + sum_i tau_ik P_i'
Psi_i^-1 y_i (update_inv_VEM / update_mean_VEM, Computing_functions.R:690-731), plus scalar sums (pi_k, ELBO,
common-hp objectives). So individuals are split into contiguous blocks balanced by sum n_i^3, every rank keeps
the SAME union grid, the prior term K_k^-1 is contributed by rank 0 only, and one all-reduce(sum) of the
K*N*(N+1) accumulator doubles per E-step combines the ranks (NCCL inside the CUDA library; gloo in the CPU
tests of this host logic).
"""
from __future__ import annotations

import numpy as np


def plan(n_obs, nranks):
    """Contiguous split of individuals over ranks, balanced by sum n_i^3 (the per-individual E-step cost).
    Returns boundaries b[0..nranks] with rank r owning individuals [b[r], b[r+1])."""
    n_obs = np.asarray(n_obs, dtype=np.float64)
    M = n_obs.shape[0]
    if nranks <= 1:
        return np.array([0, M], dtype=np.int64)
    cost = np.cumsum(n_obs ** 3)
    total = cost[-1]
    b = [0]
    for r in range(1, nranks):
        cut = int(np.searchsorted(cost, total * r / nranks, side="left")) + 1
        cut = min(max(cut, b[-1] + 1), M - (nranks - r))  # every rank gets at least one individual
        b.append(cut)
    b.append(M)
    return np.array(b, dtype=np.int64)


def take(offsets, grid_idx, y, lo, hi):
    """CSR slice of individuals [lo, hi) (grid indices stay global: every rank shares the union grid)."""
    offsets = np.asarray(offsets)
    a, b = int(offsets[lo]), int(offsets[hi])
    return (offsets[lo:hi + 1] - offsets[lo]).astype(np.int32), np.asarray(grid_idx)[a:b].copy(), np.asarray(y)[a:b].copy()


def prior_scale(rank):
    """the prior precision K_k^-1 (and K_k^-1 m_k) enters the all-reduced sums exactly once"""
    return 1.0 if rank == 0 else 0.0
