"""ctypes binding of libmagmaclust_b200.so (the C ABI declared in include/mcb.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible, every compute
call raises. Build the library with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C magmaclust_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCB_LIB: another BUILD of the same CUDA library (compile-time experiment variants); never a different implementation
LIB_PATH = os.environ.get("MCB_LIB") or os.path.join(_HERE, "lib", "libmagmaclust_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)
c_ll_p = C.POINTER(C.c_longlong)
c_ubyte_p = C.POINTER(C.c_ubyte)
H = C.c_void_p
# mcb_logl_fn: int fn(void* ctx, int npts, const int* pt_ind, const double* pt_theta, double* out_logl)
LOGL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_int_p, c_double_p, c_double_p)

# name -> (restype, argtypes); must list every symbol include/mcb.h declares (checked by tests/test_abi.py)
SIGNATURES = {
    "mcb_version": (C.c_int, []),
    "mcb_device_count": (C.c_int, []),
    "mcb_create": (C.c_int, [C.POINTER(H), C.c_int]),
    "mcb_destroy": (None, [H]),
    "mcb_last_error": (C.c_char_p, [H]),
    "mcb_comm_unique_id": (C.c_int, [c_ubyte_p]),
    "mcb_comm_init": (C.c_int, [H, C.c_int, C.c_int, c_ubyte_p]),
    "mcb_set_data": (C.c_int, [H, C.c_int, C.c_int, C.c_longlong, c_int_p, c_int_p, c_double_p, c_double_p, C.c_longlong]),
    "mcb_set_state": (C.c_int, [H, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p]),
    "mcb_get_hp": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "mcb_get_tau": (C.c_int, [H, c_double_p]),
    "mcb_e_step": (C.c_int, [H]),
    "mcb_get_mean": (C.c_int, [H, c_double_p]),
    "mcb_get_cov": (C.c_int, [H, C.c_int, c_double_p]),
    "mcb_get_tau_new": (C.c_int, [H, c_double_p]),
    "mcb_get_mat_logL": (C.c_int, [H, c_double_p]),
    "mcb_get_logdet_cov": (C.c_int, [H, c_double_p]),
    "mcb_accept_tau": (C.c_int, [H]),
    "mcb_e_step_VEM": (C.c_int, [H, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p,
                                 c_double_p, c_double_p, c_double_p, c_double_p]),
    "mcb_set_posterior": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "mcb_eval_indiv": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "mcb_eval_indiv_common": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "mcb_eval_mean": (C.c_int, [H, C.c_int, c_double_p, C.c_int, C.c_double, c_double_p, c_double_p]),
    "mcb_eval_mean_common": (C.c_int, [H, c_double_p, C.c_int, C.c_double, c_double_p, c_double_p]),
    "mcb_elbo": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "mcb_bic_terms": (C.c_int, [H, c_double_p]),
    "mcb_m_step": (C.c_int, [H, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p]),
    "mcb_accept_hp": (C.c_int, [H]),
    "mcb_get_new_hp": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "mcb_vem_iteration": (C.c_int, [H, C.c_int, C.c_int, c_double_p]),
    "mcb_vem_iteration_host": (C.c_int, [H, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "mcb_posterior_mu_k": (C.c_int, [H, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "mcb_set_pred_posterior": (C.c_int, [H, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "mcb_predict": (C.c_int, [H, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p, c_double_p, C.c_int, c_int_p,
                              c_double_p, c_double_p, c_double_p]),
    "mcb_new_indiv_logl": (C.c_int, [H, C.c_int, c_int_p, c_int_p, c_double_p, C.c_int, c_int_p, c_double_p, c_double_p]),
    "mcb_train_new_indiv": (C.c_int, [H, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p, C.c_int, c_double_p, C.c_int,
                                      c_double_p, c_double_p, c_int_p]),
    "mcb_new_indiv_em": (C.c_int, [C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, LOGL_FN, C.c_void_p,
                                   c_double_p, c_double_p, c_int_p]),
    "mcb_kmeans": (C.c_int, [H, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, c_int_p, C.c_int, c_int_p, c_double_p,
                             c_double_p, c_int_p]),
    "mcb_kern_to_cov": (C.c_int, [H, C.c_int, c_double_p, c_double_p, C.c_double, c_double_p]),
    "mcb_kern_to_inv": (C.c_int, [H, C.c_int, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p]),
    "mcb_dmvnorm": (C.c_int, [H, C.c_int, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p]),
    "mcb_gp_mod": (C.c_int, [H, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p,
                             C.c_int, C.c_double, c_double_p, c_double_p]),
    "mcb_last_timings": (C.c_int, [H, c_float_p, c_ll_p]),
    "mcb_last_mstep_stats": (C.c_int, [H, c_ll_p]),
    "mcb_last_mstep_indiv": (C.c_int, [H, c_int_p, c_int_p, c_int_p]),
    "mcb_pin_host": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mcb_unpin_host": (C.c_int, [C.c_void_p]),
    "mcb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "mcb_host_free": (C.c_int, [C.c_void_p]),
    "mcb_timer_start": (C.c_int, [H]),
    "mcb_timer_stop": (C.c_int, [H, c_float_p]),
    "mcb_flush_l2": (C.c_int, [H]),
    "mcb_measure_fp64_peaks": (C.c_int, [H, c_double_p]),
}

_lib = None


class McbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmagmaclust_b200 error {code}: {msg}")
        self.code = code


def load():
    """dlopen the library and set the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library has not been built (run __graft_entry__.build()). "
            "magmaclust_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    """pointer to a C-contiguous float64 array (None -> NULL)"""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
