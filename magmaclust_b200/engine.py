"""Array-level wrapper over the C ABI: one `Engine` = one `mcb_handle` = one GPU.

This is host plumbing only (numpy arrays in, numpy arrays out); all arithmetic happens in the CUDA
library. `magmaclust_b200.magma` builds the reference's R-shaped API (dicts / named matrices) on top.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import McbError, c_double_p, dptr, f64, i32, iptr

PHASES = ("indiv_factor_scatter", "dense_hyperpost", "indiv_tau", "mstep_indiv", "mstep_mean", "elbo",
          "predict", "collectives")


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if self.lib.mcb_device_count() <= 0:
            raise RuntimeError("magmaclust_b200: no CUDA device visible and there is no CPU fallback")
        h = _lib.H()
        rc = self.lib.mcb_create(C.byref(h), device)
        if rc != 0:
            raise McbError(rc, (self.lib.mcb_last_error(None) or b"").decode())
        self.h = h
        self.M = self.N = self.K = 0
        self.T = 0
        self.Kp = 0   # clusters of the resident prediction posterior
        self._pinned_allocs = []

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            for p in self._pinned_allocs:
                self.lib.mcb_host_free(p)
            self._pinned_allocs = []
            self.lib.mcb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise McbError(rc, (self.lib.mcb_last_error(self.h) or b"").decode())

    # ------------------------------------------------------------------------------------------
    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.mcb_comm_init(self.h, nranks, rank, buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _lib.load()
        buf = (C.c_ubyte * 128)()
        rc = lib.mcb_comm_unique_id(buf)
        if rc != 0:
            raise McbError(rc, "mcb_comm_unique_id failed (NCCL not loadable?)")
        return bytes(buf)

    # ------------------------------------------------------------------------------------------
    def set_data(self, offsets, grid_idx, y, grid, M_total=None):
        offsets, grid_idx, y, grid = i32(offsets), i32(grid_idx), f64(y), f64(grid)
        M, N, P = offsets.shape[0] - 1, grid.shape[0], y.shape[0]
        self._ck(self.lib.mcb_set_data(self.h, M, N, P, iptr(offsets), iptr(grid_idx), dptr(y), dptr(grid),
                                       M if M_total is None else M_total))
        self.M, self.N, self.P = M, N, P
        self.K = 0

    def set_state(self, theta_k, theta_i, pi_k, m_k, tau):
        theta_k, theta_i, pi_k, m_k, tau = f64(theta_k), f64(theta_i), f64(pi_k), f64(m_k).ravel(), f64(tau)
        K = theta_k.shape[0]
        assert theta_k.shape == (K, 3) and theta_i.shape == (self.M, 3) and tau.shape == (self.M, K)
        self._ck(self.lib.mcb_set_state(self.h, K, dptr(theta_k), dptr(theta_i), dptr(pi_k), dptr(m_k),
                                        m_k.shape[0], dptr(tau)))
        self.K = K

    def get_hp(self):
        tk, ti, pi = np.empty((self.K, 3)), np.empty((self.M, 3)), np.empty(self.K)
        self._ck(self.lib.mcb_get_hp(self.h, dptr(tk), dptr(ti), dptr(pi)))
        return tk, ti, pi

    def get_tau(self):
        t = np.empty((self.M, self.K))
        self._ck(self.lib.mcb_get_tau(self.h, dptr(t)))
        return t

    # ------------------------------------------------------------------------------------------
    def e_step(self):
        self._ck(self.lib.mcb_e_step(self.h))

    def get_mean(self):
        m = np.empty((self.K, self.N))
        self._ck(self.lib.mcb_get_mean(self.h, dptr(m)))
        return m

    def get_cov(self, k=None):
        if k is not None:
            c = np.empty((self.N, self.N))
            self._ck(self.lib.mcb_get_cov(self.h, k, dptr(c)))
            return c
        return np.stack([self.get_cov(kk) for kk in range(self.K)])

    def get_tau_new(self):
        t = np.empty((self.M, self.K))
        self._ck(self.lib.mcb_get_tau_new(self.h, dptr(t)))
        return t

    def get_mat_logL(self):
        t = np.empty((self.M, self.K))
        self._ck(self.lib.mcb_get_mat_logL(self.h, dptr(t)))
        return t

    def get_logdet_cov(self):
        t = np.empty(self.K)
        self._ck(self.lib.mcb_get_logdet_cov(self.h, dptr(t)))
        return t

    def accept_tau(self):
        self._ck(self.lib.mcb_accept_tau(self.h))

    def e_step_VEM(self, theta_k, theta_i, pi_k, m_k, old_tau, want_cov=True, want_logL=False):
        """host buffers in, host buffers out: the call the `.Call` shim makes for e_step_VEM"""
        theta_k, theta_i, pi_k, m_k, old_tau = f64(theta_k), f64(theta_i), f64(pi_k), f64(m_k).ravel(), f64(old_tau)
        K = theta_k.shape[0]
        mean = np.empty((K, self.N))
        cov = np.empty((K, self.N, self.N)) if want_cov else None
        tau = np.empty((self.M, K))
        logL = np.empty((self.M, K)) if want_logL else None
        self._ck(self.lib.mcb_e_step_VEM(self.h, K, dptr(theta_k), dptr(theta_i), dptr(pi_k), dptr(m_k), m_k.shape[0],
                                         dptr(old_tau), dptr(mean), dptr(cov), dptr(tau), dptr(logL)))
        self.K = K
        return mean, cov, tau, logL

    def set_posterior(self, mean, cov, tau):
        mean, cov = f64(mean), f64(cov)
        tau = f64(tau) if tau is not None else None   # None: responsibilities computed on the device (update_tau_i_k_VEM)
        assert mean.shape == (self.K, self.N) and cov.shape == (self.K, self.N, self.N)
        assert tau is None or tau.shape == (self.M, self.K)
        self._ck(self.lib.mcb_set_posterior(self.h, dptr(mean), dptr(cov), dptr(tau)))

    # ------------------------------------------------------------------------------------------
    def eval_indiv(self, theta, grad=True):
        theta = f64(theta)
        assert theta.shape == (self.M, 3)
        f = np.empty(self.M)
        g = np.empty((self.M, 3)) if grad else None
        self._ck(self.lib.mcb_eval_indiv(self.h, dptr(theta), dptr(f), dptr(g)))
        return f, g

    def eval_indiv_common(self, theta, grad=True):
        theta = f64(theta)
        f = np.empty(1)
        g = np.empty(3) if grad else None
        self._ck(self.lib.mcb_eval_indiv_common(self.h, dptr(theta), dptr(f), dptr(g)))
        return float(f[0]), g

    def eval_mean(self, k, theta, pen_diag=0.0, grad=True):
        theta = f64(theta)
        nhp = theta.shape[0]
        f = np.empty(1)
        g = np.empty(nhp) if grad else None
        self._ck(self.lib.mcb_eval_mean(self.h, k, dptr(theta), nhp, float(pen_diag), dptr(f), dptr(g)))
        return float(f[0]), g

    def eval_mean_common(self, theta, pen_diag=0.0, grad=True):
        theta = f64(theta)
        nhp = theta.shape[0]
        f = np.empty(1)
        g = np.empty(nhp) if grad else None
        self._ck(self.lib.mcb_eval_mean_common(self.h, dptr(theta), nhp, float(pen_diag), dptr(f), dptr(g)))
        return float(f[0]), g

    def elbo(self, theta_k=None, theta_i=None):
        out = np.empty(2)
        tk = f64(theta_k) if theta_k is not None else None
        ti = f64(theta_i) if theta_i is not None else None
        self._ck(self.lib.mcb_elbo(self.h, dptr(tk), dptr(ti), dptr(out)))
        return float(out[0]), float(out[1])

    def bic_terms(self):
        out = np.empty(3)
        self._ck(self.lib.mcb_bic_terms(self.h, dptr(out)))
        return float(out[0]), float(out[1]), float(out[2])

    def m_step(self, common_hp_k=True, common_hp_i=True):
        tk, ti, pi = np.empty((self.K, 3)), np.empty((self.M, 3)), np.empty(self.K)
        stats = np.zeros(4, dtype=np.int32)
        self._ck(self.lib.mcb_m_step(self.h, int(common_hp_k), int(common_hp_i), dptr(tk), dptr(ti), dptr(pi),
                                     iptr(stats)))
        return tk, ti, pi, {"nfev_i": int(stats[0]), "nfev_k": int(stats[1]), "nit_i": int(stats[2]),
                            "nit_k": int(stats[3])}

    def get_new_hp(self):
        tk, ti, pi = np.empty((self.K, 3)), np.empty((self.M, 3)), np.empty(self.K)
        self._ck(self.lib.mcb_get_new_hp(self.h, dptr(tk), dptr(ti), dptr(pi)))
        return tk, ti, pi

    def accept_hp(self):
        self._ck(self.lib.mcb_accept_hp(self.h))

    def vem_iteration(self, common_hp_k=True, common_hp_i=True):
        out = np.empty(3)
        self._ck(self.lib.mcb_vem_iteration(self.h, int(common_hp_k), int(common_hp_i), dptr(out)))
        self.last_vem_status = int(out[2])   # 0 go on, 1 converged, 2 M-step failure (stop without appending, MC.R:62-67)
        return float(out[0]), float(out[1]), bool(out[2] != 0.0)

    def vem_iteration_host(self, common_hp_k, common_hp_i, out_mean=None, out_cov=None, out_tau=None):
        """the same iteration with `param` of its E-step delivered to the given host arrays ([K][N], [K][N][N], [M][K],
        C-contiguous float64; pinned ones make the copies overlap the M-step)"""
        out = np.empty(3)
        for a, shp in ((out_mean, (self.K, self.N)), (out_cov, (self.K, self.N, self.N)), (out_tau, (self.M, self.K))):
            assert a is None or (a.dtype == np.float64 and a.flags.c_contiguous and a.shape == shp)
        self._ck(self.lib.mcb_vem_iteration_host(self.h, int(common_hp_k), int(common_hp_i), dptr(out), dptr(out_mean),
                                                 dptr(out_cov), dptr(out_tau)))
        self.last_vem_status = int(out[2])
        return float(out[0]), float(out[1]), bool(out[2] != 0.0)

    # ------------------------------------------------------------------------------------------
    def kern_to_cov(self, x, theta2, sigma):
        """vector branch of kern_to_cov (CF.R:68-86) for SORTED timestamps x; n x n array"""
        x, th = f64(x), f64(theta2)
        out = np.empty((x.shape[0], x.shape[0]))
        self._ck(self.lib.mcb_kern_to_cov(self.h, x.shape[0], dptr(x), dptr(th), float(sigma), dptr(out)))
        return out

    def kern_to_inv(self, x, theta2, sigma):
        """vector branch of kern_to_inv (CF.R:130-149) for SORTED timestamps x; (inverse, log det of the covariance)"""
        x, th = f64(x), f64(theta2)
        out = np.empty((x.shape[0], x.shape[0]))
        ld = C.c_double(0.0)
        self._ck(self.lib.mcb_kern_to_inv(self.h, x.shape[0], dptr(x), dptr(th), float(sigma), dptr(out), C.byref(ld)))
        return out, float(ld.value)

    def dmvnorm(self, x, mu, inv_Sigma, log=False):
        """dmvnorm (CF.R:10-35) from a precision matrix"""
        x, mu, P = f64(np.atleast_1d(x)), f64(np.atleast_1d(mu)), f64(np.atleast_2d(inv_Sigma))
        n = mu.shape[0]
        if P.shape != (n, n) or x.shape[0] != n:
            raise ValueError("incompatible arguments")
        out = C.c_double(0.0)
        self._ck(self.lib.mcb_dmvnorm(self.h, n, dptr(x), dptr(mu), dptr(P), 1 if log else 0, C.byref(out)))
        return float(out.value)

    # ------------------------------------------------------------------------------------------
    def posterior_mu_k(self, timestamps, tau=None, want_mean=True, want_cov=True):
        ts = f64(timestamps)
        T = ts.shape[0]
        mean = np.empty((self.K, T)) if want_mean else None
        cov = np.empty((self.K, T, T)) if want_cov else None
        t = f64(tau) if tau is not None else None
        self._ck(self.lib.mcb_posterior_mu_k(self.h, T, dptr(ts), dptr(t), dptr(mean), dptr(cov)))
        self.T = T
        self.Kp = self.K
        return mean, cov

    def set_pred_posterior(self, timestamps, mean, cov, pi_k=None):
        """install an explicit prediction posterior (list_mu of pred_gp_clust, MC.R:235) instead of the result of
        posterior_mu_k: timestamps [T] sorted, mean [K][T], cov [K][T][T], pi_k [K] (None: 1/K)"""
        ts, mean, cov = f64(timestamps), f64(mean), f64(cov)
        K, T = mean.shape
        assert ts.shape == (T,) and cov.shape == (K, T, T)
        pi = f64(pi_k) if pi_k is not None else None
        self._ck(self.lib.mcb_set_pred_posterior(self.h, K, T, dptr(ts), dptr(mean), dptr(cov), dptr(pi)))
        self.T, self.Kp = T, K

    def gp_mod(self, t, y, mean, new_cov, theta, pen_diag=0.0, grad=True):
        """logL_GP_mod / gr_GP_mod (Z = 1) or the common_hp_k forms (Z clusters) on explicit arguments (CF.R:194-216, 250-278,
        448-475, 510-537): t [n], y [Z][n], mean [Z] or [Z][n], new_cov [Z][n][n], theta [2 or 3] -> (f, g or None)"""
        t, y, mean, new_cov, theta = f64(t), f64(np.atleast_2d(y)), f64(mean).ravel(), f64(new_cov), f64(theta)
        Z, n = y.shape
        new_cov = new_cov.reshape(Z, n, n)
        nhp = theta.shape[0]
        f = np.empty(1)
        g = np.empty(nhp) if grad else None
        self._ck(self.lib.mcb_gp_mod(self.h, n, Z, dptr(t), dptr(y), dptr(mean), mean.shape[0], dptr(new_cov), dptr(theta),
                                     nhp, float(pen_diag), dptr(f), dptr(g)))
        return float(f[0]), g

    def pinned_empty(self, shape):
        """float64 array in page-locked host memory (mcb_host_alloc), released with the engine: destination for results that
        cross PCIe on every call (predict(out=...))"""
        n = int(np.prod(shape))
        p = C.c_void_p()
        self._ck(self.lib.mcb_host_alloc(C.byref(p), max(8, 8 * n)))
        self._pinned_allocs.append(p)
        return np.ctypeslib.as_array(C.cast(p, c_double_p), shape=(max(1, n),))[:n].reshape(shape)

    def predict(self, offsets, obs_idx, y, theta_new, pred_idx=None, pi_k=None, out=None):
        """out = (tau [B][K], mean [B][K][Tp], var [B][K][Tp]): caller-owned result buffers (reuse them across calls; page-locked
        ones from pinned_empty() make the copies run at PCIe rate)"""
        offsets, obs_idx, y, theta_new = i32(offsets), i32(obs_idx), f64(y), f64(theta_new)
        B = offsets.shape[0] - 1
        if out is not None:
            tau, mean, var = out
            pred_idx = i32(pred_idx)
            Tp = pred_idx.shape[0]
            assert tau.shape == (B, self.Kp) and mean.shape == (B, self.Kp, Tp) and var.shape == mean.shape
            pi = f64(pi_k) if pi_k is not None else None
            self._ck(self.lib.mcb_predict(self.h, B, iptr(offsets), iptr(obs_idx), dptr(y), dptr(theta_new), dptr(pi), Tp,
                                          iptr(pred_idx), dptr(tau), dptr(mean), dptr(var)))
            return tau, mean, var
        tau = np.empty((B, self.Kp))
        pi = f64(pi_k) if pi_k is not None else None
        if pred_idx is None:
            self._ck(self.lib.mcb_predict(self.h, B, iptr(offsets), iptr(obs_idx), dptr(y), dptr(theta_new), dptr(pi), 0,
                                          None, dptr(tau), None, None))
            return tau, None, None
        pred_idx = i32(pred_idx)
        Tp = pred_idx.shape[0]
        mean = np.empty((B, self.Kp, Tp))
        var = np.empty((B, self.Kp, Tp))
        self._ck(self.lib.mcb_predict(self.h, B, iptr(offsets), iptr(obs_idx), dptr(y), dptr(theta_new), dptr(pi), Tp,
                                      iptr(pred_idx), dptr(tau), dptr(mean), dptr(var)))
        return tau, mean, var

    def new_indiv_logl(self, offsets, obs_idx, y, pt_ind, pt_theta):
        """log N(y*; m_k(t*), Psi_theta(t*) + C_k(t*, t*)) for evaluation points (individual pt_ind[p], hyper-parameters
        pt_theta[p]) against the prediction posterior -> [npts][K] (NaN where not positive definite)"""
        offsets, obs_idx, y = i32(offsets), i32(obs_idx), f64(y)
        pt_ind, pt_theta = i32(pt_ind), f64(pt_theta).reshape(-1, 3)
        assert pt_ind.shape[0] == pt_theta.shape[0]
        out = np.empty((pt_ind.shape[0], self.Kp))
        self._ck(self.lib.mcb_new_indiv_logl(self.h, offsets.shape[0] - 1, iptr(offsets), iptr(obs_idx), dptr(y),
                                             pt_ind.shape[0], iptr(pt_ind), dptr(pt_theta), dptr(out)))
        return out

    def train_new_indiv(self, offsets, obs_idx, y, theta0, pi_k=None, n_loop_max=25):
        """free-hp branch of train_new_gp_EM (MC.R:142-186, repaired), batched: (theta_new [B][3], tau [B][K], loops [B])"""
        offsets, obs_idx, y, theta0 = i32(offsets), i32(obs_idx), f64(y), f64(theta0)
        B = offsets.shape[0] - 1
        per = theta0.ndim == 2
        assert theta0.shape == ((B, 3) if per else (3,))
        pi = f64(pi_k) if pi_k is not None else None
        th, tau, loops = np.empty((B, 3)), np.empty((B, self.Kp)), np.empty(B, dtype=np.int32)
        self._ck(self.lib.mcb_train_new_indiv(self.h, B, iptr(offsets), iptr(obs_idx), dptr(y), dptr(theta0), int(per),
                                              dptr(pi), int(n_loop_max), dptr(th), dptr(tau), iptr(loops)))
        return th, tau, loops

    # ------------------------------------------------------------------------------------------
    def kmeans(self, feat, K, start_rows, max_iter=100):
        """ini_kmeans' clustering step (MAGMAclust.R:585) on the device: feat [M][D], start_rows [nstart][K] (initial centres
        = those rows) -> (labels [M] in [0, K), centres [K][D], total within-cluster SS, assignment passes)"""
        feat, start_rows = f64(feat), i32(start_rows)
        M, D = feat.shape
        nstart = start_rows.shape[0]
        assert start_rows.shape == (nstart, K)
        lab, cent = np.empty(M, dtype=np.int32), np.empty((K, D))
        cost, iters = C.c_double(0.0), C.c_int(0)
        self._ck(self.lib.mcb_kmeans(self.h, M, D, dptr(feat), K, nstart, iptr(start_rows), int(max_iter), iptr(lab),
                                     dptr(cent), C.byref(cost), C.byref(iters)))
        return lab, cent, float(cost.value), int(iters.value)

    def last_timings(self):
        ms = (C.c_float * 8)()
        n = C.c_longlong(0)
        self._ck(self.lib.mcb_last_timings(self.h, ms, C.byref(n)))
        return {p: float(ms[i]) for i, p in enumerate(PHASES)}, int(n.value)

    def last_mstep_stats(self):
        out = (C.c_longlong * 6)()
        self._ck(self.lib.mcb_last_mstep_stats(self.h, out))
        return {"nfev_i_max": out[0], "nfev_k": out[1], "nit_i": out[2], "nit_k": out[3], "nfev_i_sum": out[4],
                "M_local": out[5]}

    def last_mstep_indiv(self):
        """(nfev [M], niter [M], status [M]) of the last per-individual M-step (status: 1 converged, 2 maxit, 3 abnormal line
        search, 4 zero gradient, 5 not finite at the start)"""
        a, b, c = (np.empty(self.M, dtype=np.int32) for _ in range(3))
        self._ck(self.lib.mcb_last_mstep_indiv(self.h, iptr(a), iptr(b), iptr(c)))
        return a, b, c

    def pin(self, arr):
        """page-lock a numpy array in place (cudaHostRegister); returns True on success"""
        return self.lib.mcb_pin_host(arr.ctypes.data_as(C.c_void_p), arr.nbytes) == 0

    def unpin(self, arr):
        self.lib.mcb_unpin_host(arr.ctypes.data_as(C.c_void_p))

    def timer_start(self):
        self._ck(self.lib.mcb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self.lib.mcb_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        self._ck(self.lib.mcb_flush_l2(self.h))

    def measure_fp64_peaks(self):
        out = np.empty(2)
        self._ck(self.lib.mcb_measure_fp64_peaks(self.h, dptr(out)))
        return {"dfma_tflops": float(out[0]), "dmma_tflops": float(out[1])}
