"""Host-side mirror of the reference's R interface for the VEM hot path, on top of the CUDA library.

Same function names, argument order and nested-list result shapes as `/root/reference/
Computing_functions.R` (CF.R) and `MAGMAclust.R` (MC.R), with R lists rendered as dicts:

    db        {'ID': array, 'Timestamp': array, 'Output': array}          (tibble columns, MC.R:22)
    m_k       {'K1': 0.0, ...}                                            (prior means, MC.R:16)
    hp        {'theta_k': {k: [a, b, sigma]}, 'theta_i': {id: [a, b, sigma]}, 'pi_k': array[K]}
    tau_i_k   {k: {id: tau}}                                              (CF.R:757-761)
    param     {'mean': {k: {'Timestamp', 'Output'}}, 'cov': {k: NamedMat}, 'tau_i_k': ...}   (CF.R:628)

Everything numerical runs on the GPU through `Engine` (the C ABI); this module only reshapes. Kernels are
passed as the sentinels `kernel` / `kernel_mu` (both the squared-exponential of CF.R:43-59): the reference
hard-codes SE derivatives (CF.R:436-446), so any other closure is rejected rather than silently ignored.
R itself is not available in the build image; the `.Call` shim that exposes the same C ABI to R is in
`r/` (source only) and INTEGRATION.md.
"""
from __future__ import annotations

import hashlib
import time

import numpy as np

from .engine import Engine


# ---------------------------------------------------------------------------------------------
def kernel(mat, theta=(1.0, 0.5)):
    """CF.R:43-50 — the SE kernel on squared distances (host formula kept for documentation/tests;
    the GPU path recognises this function object and evaluates the same formula on the device)."""
    return np.exp(theta[0] - np.exp(-theta[1]) / 2.0 * np.asarray(mat, dtype=np.float64))


def kernel_mu(mat, theta=(1.0, 0.5)):
    """CF.R:52-59 — same SE kernel, used for the mean processes."""
    return np.exp(theta[0] - np.exp(-theta[1]) / 2.0 * np.asarray(mat, dtype=np.float64))


def _check_kernels(*kerns):
    for k in kerns:
        if k is not kernel and k is not kernel_mu:
            raise ValueError("only the reference's squared-exponential kernels (kernel, kernel_mu) are "
                             "implemented on the GPU path; other closures are rejected (no fallback)")


def _db_digest(db):
    """content hash of a `db` (ID, Timestamp, Output columns in row order): decides whether the resident CSR copy of the
    training set is still the caller's data. O(P) host work, negligible next to an upload."""
    ids = np.asarray(db["ID"])
    h = hashlib.blake2b(digest_size=16)
    if ids.dtype.kind in "iufb":
        h.update(ids.dtype.str.encode()); h.update(np.ascontiguousarray(ids).tobytes())
    else:
        h.update("\x00".join(str(v) for v in ids.tolist()).encode())
    h.update(np.ascontiguousarray(db["Timestamp"], dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(db["Output"], dtype=np.float64).tobytes())
    return h.digest()


class NamedMat:
    """Square matrix named by timestamps, standing in for R's dimnames 'X<t>' (CF.R:83,146)."""

    __slots__ = ("m", "t", "_pos")

    def __init__(self, m, t):
        self.m = np.asarray(m, dtype=np.float64)
        self.t = np.asarray(t, dtype=np.float64)
        self._pos = None

    def pos(self, ts):
        if self._pos is None:
            self._pos = {float(v): j for j, v in enumerate(self.t)}
        return np.array([self._pos[float(v)] for v in np.atleast_1d(ts)], dtype=np.int64)

    def sub(self, rows, cols=None):
        r = self.pos(rows)
        c = r if cols is None else self.pos(cols)
        return self.m[np.ix_(r, c)]


# ---------------------------------------------------------------------------------------------
class Canon:
    """CSR canonicalisation of a `db` (replaces tibble filtering and string-named indexing)."""

    def __init__(self, db):
        ids = np.asarray(db["ID"])
        ts = np.asarray(db["Timestamp"], dtype=np.float64)
        ys = np.asarray(db["Output"], dtype=np.float64)
        _, first = np.unique(ids, return_index=True)
        self.ids = [i.item() if isinstance(i, np.generic) else i for i in ids[np.sort(first)]]  # unique(db$ID)
        order = np.concatenate([np.nonzero(ids == i)[0] for i in self.ids]) if len(self.ids) else np.zeros(0, int)
        counts = np.array([(ids == i).sum() for i in self.ids], dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.grid = np.unique(ts)  # sort(unique(db$Timestamp)), CF.R:601
        self.t = ts[order]
        self.y = np.ascontiguousarray(ys[order])
        self.grid_idx = np.searchsorted(self.grid, self.t).astype(np.int32)
        for a in range(len(self.ids)):
            seg = self.grid_idx[self.offsets[a]:self.offsets[a + 1]]
            if np.any(np.diff(seg) <= 0):
                raise ValueError(f"rows of individual {self.ids[a]!r} must be sorted by Timestamp and unique "
                                 "(the reference assumes it, SURVEY.md §8.0 Q2)")
        self.M, self.N = len(self.ids), self.grid.shape[0]
        self.fingerprint = _db_digest(db)


class Session:
    """One GPU engine plus the bookkeeping that lets R-style stateless calls reuse resident device state."""

    def __init__(self, device: int = 0):
        self.eng = Engine(device)
        self.canon = None
        self._token = None
        self._post_ref = None   # (mean dict, cov dict) objects of the resident hyper-posterior (strong references)
        self._epoch = 0
        self._pred_ref = None   # (mean dict, cov dict) objects of the resident PREDICTION posterior
        self._mk = None

    def close(self):
        self.eng.close()

    @property
    def _post_token(self):
        return self._token

    @_post_token.setter
    def _post_token(self, v):
        self._token = v
        if v is None:
            self._post_ref = None   # the resident hyper-posterior no longer is the one those objects describe

    def set_state(self, theta_k, theta_i, pi, mk, tau):
        self.eng.set_state(theta_k, theta_i, pi, mk, tau)
        self._mk = np.asarray(mk, dtype=np.float64).ravel().copy()   # resident prior means (K or K*N values)

    # -- data ------------------------------------------------------------------------------------
    def ensure_data(self, db):
        if self.canon is None or _db_digest(db) != self.canon.fingerprint:
            c = Canon(db)
            self.eng.set_data(c.offsets, c.grid_idx, c.y, c.grid)
            self.canon = c
            self._post_token = None
        return self.canon

    # -- packing -----------------------------------------------------------------------------------
    def pack_hp(self, hp, names_k):
        c = self.canon
        theta_k = np.array([np.asarray(hp["theta_k"][k], dtype=np.float64)[:3] for k in names_k])
        theta_i = np.array([np.asarray(hp["theta_i"][i], dtype=np.float64).ravel()[:3] for i in c.ids])
        return theta_k, theta_i

    def pack_tau(self, tau_i_k, names_k):
        c = self.canon
        return np.array([[tau_i_k[k][i] for k in names_k] for i in c.ids], dtype=np.float64)

    def unpack_tau(self, arr, names_k):
        c = self.canon
        return {k: {i: float(arr[r, ck]) for r, i in enumerate(c.ids)} for ck, k in enumerate(names_k)}

    def pack_mk(self, m_k, names_k):
        vals = [np.atleast_1d(np.asarray(m_k[k], dtype=np.float64)) for k in names_k]
        if all(v.shape[0] == 1 for v in vals):
            return np.array([v[0] for v in vals])
        N = self.canon.N
        return np.concatenate([np.repeat(v, N) if v.shape[0] == 1 else v for v in vals])

    def param_from_device(self, names_k, tau_arr, want_cov=True):
        c = self.canon
        mean = self.eng.get_mean()
        out = {"mean": {k: {"Timestamp": c.grid, "Output": mean[ck]} for ck, k in enumerate(names_k)},
               "tau_i_k": self.unpack_tau(tau_arr, names_k)}
        if want_cov:
            out["cov"] = {k: NamedMat(self.eng.get_cov(ck), c.grid) for ck, k in enumerate(names_k)}
        self._epoch += 1
        self._post_token = (id(self), self._epoch)
        out["_token"] = self._post_token
        self._post_ref = (out["mean"], out.get("cov"))
        return out

    def ensure_pred(self, mean_k, cov_k, pi_k=None):
        """make (mean_k, cov_k) the resident prediction posterior. No copies when they are the very objects the last
        posterior_mu_k of this session returned (identity, held by strong reference); anything else -- a mu_k supplied by
        the caller, read from disk, or computed by another session / an earlier call -- is uploaded."""
        ref = self._pred_ref
        if ref is not None and mean_k is ref[0] and cov_k is ref[1]:
            return
        names_k = list(mean_k.keys())
        ts = np.asarray(mean_k[names_k[0]]["Timestamp"], dtype=np.float64)
        mean = np.array([np.asarray(mean_k[k]["Output"], dtype=np.float64) for k in names_k])
        cov = np.array([np.asarray(getattr(cov_k[k], "m", cov_k[k]), dtype=np.float64) for k in names_k])
        self.eng.set_pred_posterior(ts, mean, cov, pi_k)
        self._pred_ref = (mean_k, cov_k)

    def ensure_posterior(self, hp, m_k, param):
        """make (hp, param) the resident state; no copies when `param` is what the last E-step left"""
        names_k = list(param["mean"].keys())
        if param.get("_token") is not None and param.get("_token") == self._post_token and \
                self._hp_matches(hp, names_k):
            return names_k
        theta_k, theta_i = self.pack_hp(hp, names_k)
        tau = self.pack_tau(param["tau_i_k"], names_k)
        pi = np.asarray(hp.get("pi_k", np.full(len(names_k), 1.0 / len(names_k))), dtype=np.float64)
        self.set_state(theta_k, theta_i, pi, self.pack_mk(m_k, names_k), tau)
        mean = np.array([np.asarray(param["mean"][k]["Output"], dtype=np.float64) for k in names_k])
        cov = np.array([param["cov"][k].m for k in names_k])
        self.eng.set_posterior(mean, cov, tau)
        self._post_token = None
        return names_k

    def _hp_matches(self, hp, names_k):
        tk, ti, _ = self.eng.get_hp()
        a, b = self.pack_hp(hp, names_k)
        return np.array_equal(tk, a) and np.array_equal(ti, b)


_default_session = None


def default_session():
    global _default_session
    if _default_session is None:
        _default_session = Session(0)
    return _default_session


# ---------------------------------------------------------------------------------------------
# building blocks the reference also exposes (CF.R:10-191, 436-446)
# ---------------------------------------------------------------------------------------------
def mat_dist(x, y):
    """CF.R:37-40 -- squared distances between all pairs of x and y (host: a pure index pattern, no arithmetic to offload)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    d = x[:, None] - y[None, :]
    return d * d


def deriv_hp1(mat, theta):
    """CF.R:436-439 -- d kernel / d theta1 on squared distances (host formula; on the device it is fused into the gradient kernels)."""
    return np.exp(theta[0] - np.exp(-theta[1]) / 2.0 * np.asarray(mat, dtype=np.float64))


def deriv_hp2(mat, theta):
    """CF.R:441-446 -- d kernel / d theta2 on squared distances."""
    grad = 0.5 * np.exp(-theta[1]) * np.asarray(mat, dtype=np.float64)
    return np.exp(theta[0]) * grad * np.exp(-grad)


def _per_id_theta(db, theta, sigma):
    ids = Canon(db).ids
    if isinstance(theta, dict):
        return ids, {i: np.asarray(theta[i], dtype=np.float64) for i in ids}
    th = np.r_[np.asarray(theta, dtype=np.float64)[:2], sigma]
    return ids, {i: th for i in ids}


def kern_to_cov(x, kern=kernel, theta=(1.0, 0.5), sigma=0.2, session=None):
    """CF.R:61-121 -- covariance matrix of a vector of timestamps (sorted like the reference does, CF.R:76), or one per
    individual for a `db` (theta: dict ID -> (a, b, sigma), or one (a, b) with `sigma`). Built on the GPU."""
    _check_kernels(kern)
    s = session or default_session()
    if not isinstance(x, dict):
        xs = np.sort(np.atleast_1d(np.asarray(x, dtype=np.float64)))
        return NamedMat(s.eng.kern_to_cov(xs, np.asarray(theta, dtype=np.float64)[:2], sigma), xs)
    ids, th = _per_id_theta(x, theta, sigma)
    idc, ts = np.asarray(x["ID"]), np.asarray(x["Timestamp"], dtype=np.float64)
    out = {}
    for i in ids:
        xs = np.sort(ts[idc == i])
        out[i] = NamedMat(s.eng.kern_to_cov(xs, th[i][:2], th[i][2]), xs)
    return next(iter(out.values())) if len(out) == 1 else out


def kern_to_inv(x, kern=kernel, theta=(1.0, 0.5), sigma=0.2, session=None):
    """CF.R:123-191 -- inverse covariance (same argument conventions as kern_to_cov; in the db branch theta[[i]][3] is
    the noise sd, CF.R:157-176). Raises McbError (MCB_ENOTPD) where the reference would switch to MASS::ginv."""
    _check_kernels(kern)
    s = session or default_session()
    if not isinstance(x, dict):
        xs = np.sort(np.atleast_1d(np.asarray(x, dtype=np.float64)))
        return NamedMat(s.eng.kern_to_inv(xs, np.asarray(theta, dtype=np.float64)[:2], sigma)[0], xs)
    ids, th = _per_id_theta(x, theta, sigma)
    idc, ts = np.asarray(x["ID"]), np.asarray(x["Timestamp"], dtype=np.float64)
    out = {}
    for i in ids:
        xs = np.sort(ts[idc == i])
        out[i] = NamedMat(s.eng.kern_to_inv(xs, th[i][:2], th[i][2])[0], xs)
    return next(iter(out.values())) if len(out) == 1 else out


def dmvnorm(x, mu, inv_Sigma, log=False, session=None):
    """CF.R:10-35 -- Gaussian (log-)density from a precision matrix."""
    s = session or default_session()
    P = getattr(inv_Sigma, "m", inv_Sigma)
    return s.eng.dmvnorm(x, mu, P, log)


# ---------------------------------------------------------------------------------------------
# E step / M step (CF.R:589-687)
# ---------------------------------------------------------------------------------------------
def e_step_VEM(db, m_k, kern_0, kern_i, hp, old_tau_i_k, session=None, want_cov=True):
    """CF.R:589-629 -> list(mean, cov, tau_i_k)."""
    _check_kernels(kern_0, kern_i)
    s = session or default_session()
    s.ensure_data(db)
    names_k = list(m_k.keys())
    theta_k, theta_i = s.pack_hp(hp, names_k)
    s.set_state(theta_k, theta_i, np.asarray(hp["pi_k"], dtype=np.float64), s.pack_mk(m_k, names_k),
                    s.pack_tau(old_tau_i_k, names_k))
    s.eng.e_step()
    out = s.param_from_device(names_k, s.eng.get_tau_new(), want_cov)
    out["mat_logL"] = s.eng.get_mat_logL().T  # K x M like CF.R:738
    return out


def update_tau_i_k_VEM(db, m_k, mean_k, cov_k, kern_i, hp, pi_k, session=None):
    """CF.R:733-762 -> tau_i_k (list over clusters of lists over individuals) from given means / covariances of the mean
    processes: log-likelihoods, max-shift and normalisation on the device (pi_k multiplied positionally, quirk Q3)."""
    _check_kernels(kern_i)
    s = session or default_session()
    s.ensure_data(db)
    names_k = list(m_k.keys())
    hp_full = hp if "theta_k" in hp else dict(hp, theta_k={k: np.array([1.0, 1.0, 0.2]) for k in names_k})  # unused here
    theta_k, theta_i = s.pack_hp(hp_full, names_k)
    M = theta_i.shape[0]
    s.set_state(theta_k, theta_i, np.asarray(pi_k, dtype=np.float64), s.pack_mk(m_k, names_k),
                    np.full((M, len(names_k)), 1.0 / len(names_k)))
    mean = np.array([np.asarray(mean_k[k]["Output"], dtype=np.float64) for k in names_k])
    cov = np.array([np.asarray(getattr(cov_k[k], "m", cov_k[k]), dtype=np.float64) for k in names_k])
    s.eng.set_posterior(mean, cov, None)
    s._post_token = None
    return s.unpack_tau(s.eng.get_tau_new(), names_k)


def m_step_VEM(db, old_hp, list_mu_param, kern_0, kern_i, m_k, common_hp_k, common_hp_i, session=None, stats=None):
    """CF.R:631-687 -> list(theta_k, theta_i, pi_k); optimiser = the library's L-BFGS (csrc/lbfgs.h)."""
    _check_kernels(kern_0, kern_i)
    s = session or default_session()
    c = s.ensure_data(db)
    names_k = s.ensure_posterior(old_hp, m_k, list_mu_param)
    tk, ti, pi, st = s.eng.m_step(common_hp_k, common_hp_i)
    if stats is not None:
        stats.update(st)
    return {"theta_k": {k: tk[ck] for ck, k in enumerate(names_k)},
            "theta_i": {i: ti[r] for r, i in enumerate(c.ids)}, "pi_k": pi}


# ---------------------------------------------------------------------------------------------
# optimiser callbacks (CF.R:194-325, 448-586) — each is one evaluation on the device
# ---------------------------------------------------------------------------------------------
def _indiv_rows(s, hp3, i):
    c = s.canon
    _, ti, _ = s.eng.get_hp()
    ti = ti.copy()
    ti[c.ids.index(i)] = hp3
    return ti, c.ids.index(i)


def _require_token(s, mu_k_param, what):
    if mu_k_param.get("_token") is None or mu_k_param.get("_token") != s._post_token:
        raise ValueError(f"{what}: pass the param object returned by the last e_step_VEM of this session, "
                         "or call Session.ensure_posterior(hp, m_k, param) first")


def logL_clust_multi_GP(hp, db, mu_k_param, kern, session=None):
    """CF.R:218-248 for the single individual in `db` (3 hyper-parameters)."""
    _check_kernels(kern)
    s = session or default_session()
    _require_token(s, mu_k_param, "logL_clust_multi_GP")
    i = np.asarray(db["ID"])[0]
    i = i.item() if isinstance(i, np.generic) else i
    th, r = _indiv_rows(s, np.asarray(hp, dtype=np.float64)[:3], i)
    f, _ = s.eng.eval_indiv(th, grad=False)
    return float(f[r])


def gr_clust_multi_GP(hp, db, mu_k_param, kern, session=None):
    """CF.R:477-508."""
    _check_kernels(kern)
    s = session or default_session()
    _require_token(s, mu_k_param, "gr_clust_multi_GP")
    i = np.asarray(db["ID"])[0]
    i = i.item() if isinstance(i, np.generic) else i
    th, r = _indiv_rows(s, np.asarray(hp, dtype=np.float64)[:3], i)
    _, g = s.eng.eval_indiv(th, grad=True)
    return g[r]


def logL_clust_multi_GP_common_hp_i(hp, db, mu_k_param, kern, session=None):
    """CF.R:280-325."""
    _check_kernels(kern)
    s = session or default_session()
    _require_token(s, mu_k_param, "logL_clust_multi_GP_common_hp_i")
    return s.eng.eval_indiv_common(np.asarray(hp, dtype=np.float64)[:3], grad=False)[0]


def gr_clust_multi_GP_common_hp_i(hp, db, mu_k_param, kern, session=None):
    """CF.R:539-586."""
    _check_kernels(kern)
    s = session or default_session()
    _require_token(s, mu_k_param, "gr_clust_multi_GP_common_hp_i")
    return s.eng.eval_indiv_common(np.asarray(hp, dtype=np.float64)[:3], grad=True)[1]


def _mat(x):
    return np.asarray(getattr(x, "m", x), dtype=np.float64)


def _resident_cluster(s, db, mean, new_cov):
    """position of the cluster whose resident hyper-posterior mean / covariance ARE `db` / `new_cov` (object identity with
    what the last e_step_VEM / training_VEM of this session returned) and whose resident prior mean equals `mean`;
    None otherwise (the evaluation then runs on the explicit arguments)"""
    ref = s._post_ref
    if ref is None or s._post_token is None or ref[1] is None or s._mk is None:
        return None
    for ck, k in enumerate(ref[0]):
        if db is ref[0][k] and new_cov is ref[1][k]:
            mk = s._mk[ck] if s._mk.shape[0] == len(ref[0]) else s._mk[ck * s.canon.N:(ck + 1) * s.canon.N]
            m = np.atleast_1d(np.asarray(mean, dtype=np.float64))
            return ck if (m.shape[0] in (1, np.size(mk)) and np.all(m == mk)) else None
    return None


def _gp_mod(hp, db, mean, kern, new_cov, pen_diag, grad, session):
    _check_kernels(kern)
    s = session or default_session()
    hp = np.asarray(hp, dtype=np.float64).ravel()[:3]
    pen = 0.0 if pen_diag is None else float(pen_diag)
    k = _resident_cluster(s, db, mean, new_cov)
    if k is not None:
        return s.eng.eval_mean(k, hp, pen, grad=grad)
    t = np.asarray(db["Timestamp"], dtype=np.float64)
    y = np.asarray(db["Output"], dtype=np.float64)
    m = np.atleast_1d(np.asarray(mean, dtype=np.float64))
    if t.shape[0] == 1 and grad:
        # CF.R:460-463: one timestamp only -- the reference takes a branch of its own; same arithmetic on 1 x 1 matrices
        pass
    return s.eng.gp_mod(t, y[None, :], m, _mat(new_cov)[None], hp, pen, grad=grad)


def logL_GP_mod(hp, db, mean, kern, new_cov, pen_diag=None, session=None):
    """CF.R:194-216 with the reference's argument list: hp (a, b[, sigma]), db = {'Timestamp', 'Output'}, mean (scalar or
    vector), kern, new_cov (n x n), pen_diag. When (db, new_cov) are a mean process of the resident hyper-posterior the
    evaluation runs against the device-resident copies (what m_step_VEM does, CF.R:675); otherwise the arguments are
    uploaded (`mcb_gp_mod`)."""
    return _gp_mod(hp, db, mean, kern, new_cov, pen_diag, False, session)[0]


def gr_GP_mod(hp, db, mean, kern, new_cov, pen_diag=None, session=None):
    """CF.R:448-475 -> numeric[length(hp)]."""
    return _gp_mod(hp, db, mean, kern, new_cov, pen_diag, True, session)[1]


def _gp_mod_common(hp, db, mean, kern, new_cov, pen_diag, grad, session):
    _check_kernels(kern)
    s = session or default_session()
    hp = np.asarray(hp, dtype=np.float64).ravel()[:3]
    pen = 0.0 if pen_diag is None else float(pen_diag)
    names_k = list(db.keys())
    ref = s._post_ref
    if ref is not None and s._post_token is not None and db is ref[0] and new_cov is ref[1] and s._mk is not None and \
            all(_resident_cluster(s, db[k], mean[k], new_cov[k]) == ck for ck, k in enumerate(names_k)):
        return s.eng.eval_mean_common(hp, pen, grad=grad)
    t = np.asarray(db[names_k[0]]["Timestamp"], dtype=np.float64)       # t_k = db[[1]]$Timestamp, CF.R:264
    y = np.array([np.asarray(db[k]["Output"], dtype=np.float64) for k in names_k])
    ms = [np.atleast_1d(np.asarray(mean[k], dtype=np.float64)) for k in names_k]
    m = np.array([v[0] for v in ms]) if all(v.shape[0] == 1 for v in ms) else \
        np.concatenate([np.repeat(v, t.shape[0]) if v.shape[0] == 1 else v for v in ms])
    cov = np.array([_mat(new_cov[k]) for k in names_k])
    return s.eng.gp_mod(t, y, m, cov, hp, pen, grad=grad)


def logL_GP_mod_common_hp_k(hp, db, mean, kern, new_cov, pen_diag=None, session=None):
    """CF.R:250-278: db = {k: {'Timestamp', 'Output'}} (the K mean processes), mean = {k: m_k}, new_cov = {k: cov_k}."""
    return _gp_mod_common(hp, db, mean, kern, new_cov, pen_diag, False, session)[0]


def gr_GP_mod_common_hp_k(hp, db, mean, kern, new_cov, pen_diag=None, session=None):
    """CF.R:510-537."""
    return _gp_mod_common(hp, db, mean, kern, new_cov, pen_diag, True, session)[1]


def logL_monitoring_VEM(hp, db, kern_i, kern_0, mu_k_param, m_k, session=None):
    """CF.R:327-352 at `hp` against the resident hyper-posterior."""
    _check_kernels(kern_i, kern_0)
    s = session or default_session()
    s.ensure_data(db)
    _require_token(s, mu_k_param, "logL_monitoring_VEM")
    names_k = list(m_k.keys())
    theta_k, theta_i = s.pack_hp(hp, names_k)
    return s.eng.elbo(theta_k, theta_i)[0]


def BIC(hp, db, kern_0, kern_i, m_k, mu_k_param, clust=True, session=None):
    """CF.R:354-432 (clust = TRUE). The sums run on the device; the penalty counts distinct hp VALUES like the
    reference's `n_distinct(unlist(.))` (CF.R:359-361)."""
    if not clust:
        raise NotImplementedError("clust = FALSE needs the single-cluster MAGMA model of the sibling repository")
    _check_kernels(kern_0, kern_i)
    s = session or default_session()
    c = s.ensure_data(db)
    names_k = list(mu_k_param["mean"].keys())
    # evaluate at (hp, mu_k_param): install both (mat_logL is recomputed at these hp with tau frozen)
    theta_k, theta_i = s.pack_hp(hp, names_k)
    tau = s.pack_tau(mu_k_param["tau_i_k"], names_k)
    s.set_state(theta_k, theta_i, np.asarray(hp["pi_k"], dtype=np.float64), s.pack_mk(m_k, names_k), tau)
    mean = np.array([np.asarray(mu_k_param["mean"][k]["Output"], dtype=np.float64) for k in names_k])
    cov = np.array([mu_k_param["cov"][k].m for k in names_k])
    s.eng.set_posterior(mean, cov, tau)
    s._post_token = None
    ll_i, ll_k, logdet = s.eng.bic_terms()
    K, N = len(names_k), c.N
    nb_hp_i = len(np.unique(theta_i))
    nb_hp_k = len(np.unique(theta_k))
    nb_clust = len(np.unique(np.asarray(hp["pi_k"], dtype=np.float64)))
    pen = 0.5 * (nb_hp_i + nb_hp_k + nb_clust - 1) * np.log(c.M)
    return ll_i + ll_k + 0.5 * (K * (N + N * np.log(2 * np.pi)) + logdet) - pen


def model_selection(db, k_grid, ini_tau_by_k, ini_hp=None, kern_0=kernel_mu, kern_i=kernel, common_hp_k=True,
                    common_hp_i=True, session=None):
    """MC.R:101-135 for K >= 2: train one model per K, score it with BIC, keep the arg max. `ini_tau_by_k[K]` supplies
    the initial responsibilities (the reference's default is an unseeded k-means); K = 1 would need `training` from
    the sibling MAGMA repository, which the reference does not ship."""
    s = session or default_session()
    res = {}
    for K in k_grid:
        if K < 2:
            raise NotImplementedError("K = 1 needs the sibling repository's training()")
        prior_mean_k = {f"K{k + 1}": 0.0 for k in range(K)}
        model = training_VEM(db, prior_mean_k, ini_hp, kern_0, kern_i, ini_tau_by_k[K], common_hp_k, common_hp_i,
                             session=s)
        model["BIC"] = BIC(model["hp"], db, kern_0, kern_i, prior_mean_k, model["param"], True, session=s)
        res[f"K = {K}"] = model
    bics = {K: res[f"K = {K}"]["BIC"] for K in k_grid}
    res["K_max_BIC"] = max(bics, key=bics.get)
    return res


# ---------------------------------------------------------------------------------------------
# drivers (MC.R)
# ---------------------------------------------------------------------------------------------
def training_VEM(db, prior_mean_k, ini_hp=None, kern_0=kernel_mu, kern_i=kernel, ini_tau_i_k=None,
                 common_hp_k=True, common_hp_i=True, n_loop_max=15, session=None, verbose=False):
    """MC.R:16-99. The whole loop runs against device-resident state (one `mcb_vem_iteration` per pass);
    host copies happen once at the end. `ini_tau_i_k` is required (the reference's default is an unseeded
    stats::kmeans, MC.R:37, which is outside this path)."""
    _check_kernels(kern_0, kern_i)
    if ini_tau_i_k is None:
        raise ValueError("ini_tau_i_k is required")
    if ini_hp is None:
        ini_hp = {"theta_k": np.array([1.0, 1.0, 0.2]), "theta_i": np.array([1.0, 1.0, 0.2])}
    s = session or default_session()
    c = s.ensure_data(db)
    names_k = list(prior_mean_k.keys())
    K = len(names_k)
    missing = [k for k in names_k if k not in ini_tau_i_k]
    if missing or len(ini_tau_i_k) != K:
        raise ValueError(f"ini_tau_i_k must hold one entry per cluster of prior_mean_k {names_k}; got {list(ini_tau_i_k)} "
                         "(a k-means initialisation that found fewer than K clusters cannot start the VEM)")
    tau0 = s.pack_tau(ini_tau_i_k, names_k)
    theta_k = np.tile(np.asarray(ini_hp["theta_k"], dtype=np.float64)[:3], (K, 1))
    theta_i = np.tile(np.asarray(ini_hp["theta_i"], dtype=np.float64)[:3], (c.M, 1))
    # MC.R:39: sapply(tau_i_k, mean) runs in the LIST order of ini_tau_i_k (first appearance of the k-means labels,
    # MC.R:603) while update_tau_i_k_VEM multiplies pi_k positionally (CF.R:757): quirk Q3, kept
    pi = np.array([np.mean([float(v) for v in ini_tau_i_k[k].values()]) for k in ini_tau_i_k])
    t1 = time.time()
    s.set_state(theta_k, theta_i, pi, s.pack_mk(prior_mean_k, names_k), tau0)
    seq_logL = []
    cv = False
    for it in range(n_loop_max):
        elbo, eps, conv = s.eng.vem_iteration(common_hp_k, common_hp_i)
        if s.eng.last_vem_status == 2:
            # MC.R:62-67: the M-step failed -- leave the loop before logL_iterations grows, convergence stays FALSE
            if verbose:
                print(f"The M-step encountered an error at iteration : {it + 1}")
            break
        seq_logL.append(elbo)
        if verbose:
            print(it + 1, elbo, "eps", eps)
        if conv:
            cv = True
            break
    # MC.R:95 returns hp = new_hp (the last M-step's candidate) with the param of the preceding E-step (Q7)
    tk, ti, pik = s.eng.get_new_hp()
    param = s.param_from_device(names_k, s.eng.get_tau_new())
    t2 = time.time()
    hp = {"theta_k": {k: tk[ck] for ck, k in enumerate(names_k)},
          "theta_i": {i: ti[r] for r, i in enumerate(c.ids)}, "pi_k": pik}
    return {"hp": hp, "convergence": cv, "param": param, "Time_train": t2 - t1,
            "logL_iterations": np.array(seq_logL)}


def posterior_mu_k(db, timestamps, m_k, kern_0, kern_i, list_hp, session=None):
    """MC.R:196-233."""
    _check_kernels(kern_0, kern_i)
    s = session or default_session()
    c = s.ensure_data(db)
    names_k = list(list_hp["hp"]["theta_k"].keys())
    theta_k, theta_i = s.pack_hp(list_hp["hp"], names_k)
    tau = s.pack_tau(list_hp["param"]["tau_i_k"], names_k)
    pi = np.asarray(list_hp["hp"].get("pi_k", tau.mean(axis=0)), dtype=np.float64)
    tk, ti, _ = (s.eng.get_hp() if s.eng.K == len(names_k) else (None, None, None))
    if tk is None or not (np.array_equal(tk, theta_k) and np.array_equal(ti, theta_i)):
        s.set_state(theta_k, theta_i, pi, s.pack_mk(m_k, names_k), tau)
        s._post_token = None
    ts = np.asarray(timestamps, dtype=np.float64)
    mean, cov = s.eng.posterior_mu_k(ts, tau)
    out = {"mean": {k: {"Timestamp": ts, "Output": mean[ck]} for ck, k in enumerate(names_k)},
           "cov": {k: NamedMat(cov[ck], ts) for ck, k in enumerate(names_k)},
           "tau_i_k": list_hp["param"]["tau_i_k"]}
    s._pred_ref = (out["mean"], out["cov"])   # these very objects describe the resident prediction posterior
    return out


def _csr_new(new_dbs, ts):
    offs, idx, ys = [0], [], []
    for d in new_dbs:
        t = np.asarray(d["Timestamp"], dtype=np.float64)
        p = np.searchsorted(ts, t)
        if np.any(p >= ts.shape[0]) or np.any(ts[np.minimum(p, ts.shape[0] - 1)] != t):
            raise ValueError("every observation timestamp must belong to the prediction grid")
        idx.append(p)
        ys.append(np.asarray(d["Output"], dtype=np.float64))
        offs.append(offs[-1] + t.shape[0])
    return np.array(offs, np.int32), np.concatenate(idx).astype(np.int32), np.concatenate(ys)


def update_tau_star_k_EM(db, mean_k, cov_k, kern, hp, pi_k, session=None):
    """CF.R:764-786. (mean_k, cov_k) are used as given: when they are the objects the last posterior_mu_k of this session
    returned they are already resident, otherwise they are uploaded first."""
    _check_kernels(kern)
    s = session or default_session()
    s.ensure_pred(mean_k, cov_k)
    names_k = list(mean_k.keys())
    ts = np.asarray(mean_k[names_k[0]]["Timestamp"], dtype=np.float64)
    off, idx, y = _csr_new([db], ts)
    pi = np.array([pi_k[k] for k in names_k], dtype=np.float64) if isinstance(pi_k, dict) else np.asarray(pi_k)
    tau, _, _ = s.eng.predict(off, idx, y, np.asarray(hp, dtype=np.float64)[:3], None, pi)
    return {k: float(tau[0, ck]) for ck, k in enumerate(names_k)}


def train_new_gp_EM(db, param_mu_k, ini_hp=None, kern_i=kernel, hp_i=None, session=None, n_loop_max=25):
    """MC.R:137-192. hp_i given: the fixed-hp branch (MC.R:187-191). hp_i = None: the free-hp EM of MC.R:142-186 — upstream
    that branch cannot run (undefined `t`); `mcb_train_new_indiv` runs the EM it describes (E step
    update_tau_star_k_EM, M step L-BFGS on - sum_k tau_k log N(y*; m_k, Psi + C_k) with optim's numerical gradient,
    stop at 0 < sum |d theta| < 1e-3 or 25 passes). `db` may also be a LIST of new individuals: they are trained in one
    batched call and a list of results comes back."""
    pi_k = {k: float(np.mean(list(v.values()))) for k, v in param_mu_k["tau_i_k"].items()}
    if hp_i is None:
        _check_kernels(kern_i)
        s = session or default_session()
        s.ensure_pred(param_mu_k["mean"], param_mu_k["cov"])
        names_k = list(param_mu_k["mean"].keys())
        ts = np.asarray(param_mu_k["mean"][names_k[0]]["Timestamp"], dtype=np.float64)
        many = isinstance(db, (list, tuple))
        off, idx, y = _csr_new(list(db) if many else [db], ts)
        if ini_hp is None:
            ini_hp = {"theta_i": np.array([1.0, 1.0, 0.2])}
        th, tau, loops = s.eng.train_new_indiv(off, idx, y, np.asarray(ini_hp["theta_i"], dtype=np.float64).ravel()[:3],
                                               np.array([pi_k[k] for k in names_k]), n_loop_max)
        res = [{"theta_new": th[b], "tau_k": {k: float(tau[b, ck]) for ck, k in enumerate(names_k)}, "loops": int(loops[b])}
               for b in range(th.shape[0])]
        return res if many else res[0]
    new_hp = np.asarray(next(iter(hp_i["hp"]["theta_i"].values())), dtype=np.float64).ravel()[:3]
    tau_k = update_tau_star_k_EM(db, param_mu_k["mean"], param_mu_k["cov"], kern_i, new_hp, pi_k, session)
    return {"theta_new": new_hp, "tau_k": tau_k}


def pred_gp_clust(db, timestamps, list_mu, kern, hp, session=None):
    """MC.R:235-274 -> {k: {'Timestamp', 'Mean', 'Var', 'tau_k'}}."""
    _check_kernels(kern)
    s = session or default_session()
    s.ensure_pred(list_mu["mean"], list_mu["cov"])
    names_k = list(list_mu["mean"].keys())
    ts = np.asarray(list_mu["mean"][names_k[0]]["Timestamp"], dtype=np.float64)
    if timestamps is None:
        raise ValueError("timestamps must be given (and belong to the grid of list_mu)")
    tp = np.asarray(timestamps, dtype=np.float64)
    pidx = np.searchsorted(ts, tp)
    if np.any(pidx >= ts.shape[0]) or np.any(ts[np.minimum(pidx, ts.shape[0] - 1)] != tp):
        raise ValueError("prediction timestamps must belong to the grid of list_mu (MC.R:323 guarantees it)")
    off, idx, y = _csr_new([db], ts)
    _, mean, var = s.eng.predict(off, idx, y, np.asarray(hp["theta_new"], dtype=np.float64)[:3], pidx.astype(np.int32),
                                 np.full(len(names_k), 1.0 / len(names_k)))
    return {k: {"Timestamp": tp, "Mean": mean[0, ck], "Var": var[0, ck], "tau_k": hp["tau_k"][k]}
            for ck, k in enumerate(names_k)}


def full_algo_clust(db, new_db, timestamps, kern_i, ini_tau_i_k=None, common_hp_k=True, common_hp_i=True,
                    prior_mean=None, kern_0=None, list_hp=None, mu_k=None, ini_hp=None, hp_new_i=None, session=None):
    """MC.R:301-347."""
    s = session or default_session()
    if list_hp is None:
        list_hp = training_VEM(db, prior_mean, ini_hp, kern_0, kern_i, ini_tau_i_k, common_hp_k, common_hp_i,
                               session=s)
    t_pred = np.unique(np.r_[np.asarray(timestamps, dtype=np.float64), np.asarray(db["Timestamp"], dtype=np.float64),
                             np.asarray(new_db["Timestamp"], dtype=np.float64)])
    if mu_k is None:
        mu_k = posterior_mu_k(db, t_pred, prior_mean, kern_0, kern_i, list_hp, session=s)
    if hp_new_i is None:
        hp_new_i = train_new_gp_EM(new_db, mu_k, ini_hp, kern_i, hp_i=list_hp if common_hp_i else None, session=s)
    pred = pred_gp_clust(new_db, timestamps, mu_k, kern_i, hp_new_i, session=s)
    return {"Prediction": pred, "Mean_processes": mu_k,
            "Hyperparameters": {"other_i": list_hp, "new_i": hp_new_i},
            "logL_iterations": list_hp.get("logL_iterations"),
            "ari_iterations": list_hp.get("ARI_iterations")}   # MC.R:345: never set by training_VEM, i.e. NULL


from .init_io import ini_kmeans, ini_tau_i_k, read_db_csv, write_db_csv  # noqa: E402,F401  (MAGMAclust.R:562-607)


def pred_max_cluster(tau_i_k):
    """MC.R:349-355: which.max per individual (first maximum wins), 1-based."""
    names_k = list(tau_i_k.keys())
    ids = list(tau_i_k[names_k[0]].keys())
    mat = np.array([[tau_i_k[k][i] for k in names_k] for i in ids])
    return np.argmax(mat, axis=1) + 1


# ---- trained models on disk (the `.rds` lists of Simulation_study.R:554-582) --------------------------
def save_trained(res, path):
    """saveRDS(training_VEM(...), path): the R list layout of MC.R:94-98 (`rds_io.trained_to_r`)"""
    from . import rds_io
    rds_io.save_trained(res, path)


def load_trained(path_or_obj, id_type=str):
    """readRDS of a `training_VEM` result — a file, or an already decoded R list such as `[[d]]$MAGMAclust` of the study's
    stored outputs. R keeps individual IDs as names (strings); `id_type` converts them back to the type of `db['ID']`
    (e.g. `int`). Covariances come back as `NamedMat`, so the result can go straight into `posterior_mu_k` / `BIC`."""
    from . import rds_io
    obj = rds_io.read_rds(path_or_obj) if isinstance(path_or_obj, (str, bytes)) or hasattr(path_or_obj, "__fspath__") else path_or_obj
    res = rds_io.trained_from_r(obj)
    res["hp"]["theta_i"] = {id_type(i): v for i, v in res["hp"]["theta_i"].items()}
    p = res.get("param", {})
    if "tau_i_k" in p:
        p["tau_i_k"] = {k: {id_type(i): x for i, x in per_i.items()} for k, per_i in p["tau_i_k"].items()}
    if "cov" in p and "mean" in p:
        p["cov"] = {k: NamedMat(m, p["mean"][k]["Timestamp"]) for k, m in p["cov"].items()}
    return res
