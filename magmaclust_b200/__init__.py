"""magmaclust_b200 — B200-native (sm_100a) implementation of MAGMAclust's variational-EM hot path.

`magmaclust_b200.magma` mirrors the reference's R interface (e_step_VEM, m_step_VEM, training_VEM,
posterior_mu_k, pred_gp_clust, ...); `magmaclust_b200.engine.Engine` is the array-level binding of the
C ABI in include/mcb.h. All arithmetic runs in libmagmaclust_b200.so; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .engine import Engine  # noqa: F401

__all__ = ["Engine", "magma", "synth", "shard"]
