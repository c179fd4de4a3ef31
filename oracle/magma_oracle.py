"""CPU oracle for the MAGMAclust variational-EM hot path — TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's R functions (`/root/reference/
Computing_functions.R` = CF.R, `/root/reference/MAGMAclust.R` = MC.R), one Python function per R
function, same names and argument order. It exists so that the CUDA path can be checked on the
same inputs. It is NOT part of the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it. The product
(`magmaclust_b200/`) never imports it and fails loudly when its CUDA library is missing.

Parity status ("pinning"):
  * R is not installed in the build image, so the reference itself cannot be run here.
  * The oracle is pinned against the reference's own stored outputs
    (`Simulations/Training/train_for_pred_Hoo_diff_time.rds` + its input CSV; see
    `tests/golden/` and `tests/golden/make_golden.py`): `pi_k == rowMeans(tau)` bit-exactly,
    stored `tau_i_k` is a near fixed point of `e_step_VEM` (identical hard assignments in 100/100
    datasets), analytic gradients agree with central finite differences.
  * NOT pinned (no reference test or fixture constrains them): optimiser trajectories
    (`optimr::opm` -> `stats::optim` L-BFGS-B, version unpinned; here scipy's L-BFGS-B),
    `stats::kmeans` initialisation, the `MASS::ginv` fallback.

Conventions
  * `db` is a dict of equal-length arrays `{'ID', 'Timestamp', 'Output'}` (a tibble's columns).
  * R lists keyed by names become dicts in insertion order (cluster names 'K1'.., individual IDs).
  * R named matrices (`'X<t>'` dimnames, CF.R:83,146) become `NamedMat` (matrix + its timestamps);
    name lookup is by exact double equality instead of the 15-significant-digit `paste0` rendering
    (quirk Q1 of SURVEY.md §8.0).
  * Inverses/log-determinants go through LU (`numpy.linalg.inv` / `slogdet`) like R's
    `solve` / `determinant`; a `LinAlgError` falls back to `pinv` like `MASS::ginv` (CF.R:142).
"""
from __future__ import annotations

import time

import numpy as np

LOG_2PI = np.log(2.0 * np.pi)


# --------------------------------------------------------------------------------------------
# small helpers standing in for tibble / named-matrix mechanics
# --------------------------------------------------------------------------------------------
class NamedMat:
    """A square matrix whose rows/cols are named by timestamps (R: dimnames 'X<t>')."""

    __slots__ = ("m", "t", "_pos")

    def __init__(self, m, t):
        self.m = np.atleast_2d(np.asarray(m, dtype=np.float64))
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


def unique_stable(x):
    """R's unique(): first-appearance order."""
    x = np.asarray(x)
    _, first = np.unique(x, return_index=True)
    return x[np.sort(first)]


def split_db(db):
    """rows of each ID, in db order (what `db %>% filter(ID == i)` yields)."""
    ids = np.asarray(db["ID"])
    out = {}
    for i in unique_stable(ids):
        sel = np.nonzero(ids == i)[0]
        out[_key(i)] = {
            "ID": ids[sel],
            "Timestamp": np.asarray(db["Timestamp"], dtype=np.float64)[sel],
            "Output": np.asarray(db["Output"], dtype=np.float64)[sel],
        }
    return out


def _key(i):
    return i.item() if isinstance(i, np.generic) else i


def _solve(mat):
    """tryCatch(solve(mat), error = ginv(mat)) — CF.R:142,177,612."""
    try:
        return np.linalg.inv(mat)
    except np.linalg.LinAlgError:
        return np.linalg.pinv(mat)


def _dist2(x):
    """as.matrix(dist(x)^2): squared pairwise distances, zero diagonal."""
    x = np.asarray(x, dtype=np.float64)
    d = x[:, None] - x[None, :]
    return d * d


# --------------------------------------------------------------------------------------------
# CF.R:10-59 — density, distances, kernels
# --------------------------------------------------------------------------------------------
def dmvnorm(x, mu, inv_Sigma, log=False):
    """CF.R:10-35. Gaussian density from a precision matrix; log|det| by LU, sign ignored."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    n = mu.shape[0]
    inv_Sigma = np.atleast_2d(inv_Sigma)
    if inv_Sigma.shape != (n, n) or x.shape[0] != n:
        raise ValueError("incompatible arguments")
    z = x - mu
    logdetS = -np.linalg.slogdet(inv_Sigma)[1]
    ssq = z @ inv_Sigma @ z
    loglik = -(n * LOG_2PI + logdetS + ssq) / 2.0
    return loglik if log else np.exp(loglik)


def mat_dist(x, y):
    """CF.R:37-40. Squared distances between all pairs of x and y."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    d = x[:, None] - y[None, :]
    return d * d


def kernel(mat, theta=(1.0, 0.5)):
    """CF.R:43-50. Squared-exponential kernel on squared distances."""
    return np.exp(theta[0] - np.exp(-theta[1]) / 2.0 * mat)


def kernel_mu(mat, theta=(1.0, 0.5)):
    """CF.R:52-59. Same SE kernel (mean processes)."""
    return np.exp(theta[0] - np.exp(-theta[1]) / 2.0 * mat)


def deriv_hp1(mat, theta):
    """CF.R:436-439."""
    return np.exp(theta[0] - np.exp(-theta[1]) / 2.0 * mat)


def deriv_hp2(mat, theta):
    """CF.R:441-446."""
    grad = 0.5 * np.exp(-theta[1]) * mat
    return np.exp(theta[0]) * grad * np.exp(-grad)


def _cov_vector(x, kern, theta, sigma):
    """vector branch shared by kern_to_cov / kern_to_inv (CF.R:68-79, 130-141).

    `kern` is applied to a `dist` object, so `as.matrix()` leaves a zero diagonal; the diagonal is
    then `sigma^2 + exp(theta1)`. A length-1 input short-circuits to that scalar.
    """
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if x.shape[0] == 1:
        return np.array([[sigma ** 2 + np.exp(theta[0])]]), x
    x = np.sort(x)
    mat = kern(_dist2(x), theta)
    np.fill_diagonal(mat, 0.0)
    mat = mat + np.diag(np.full(x.shape[0], sigma ** 2 + np.exp(theta[0])))
    return mat, x


def kern_to_cov(x, kern=kernel, theta=(1.0, 0.5), sigma=0.2):
    """CF.R:61-121. Covariance matrix (vector input) or dict of them (db input, per-ID theta)."""
    if not isinstance(x, dict):
        mat, xs = _cov_vector(x, kern, theta, sigma)
        return NamedMat(mat, xs)
    out = {}
    for i, d in split_db(x).items():
        th = theta[i] if isinstance(theta, dict) else np.r_[np.asarray(theta, dtype=np.float64), sigma]
        mat, xs = _cov_vector(d["Timestamp"], kern, th[:2], th[2])
        out[i] = NamedMat(mat, xs)
    return next(iter(out.values())) if len(out) == 1 else out


def kern_to_inv(x, kern=kernel, theta=(1.0, 0.5), sigma=0.2):
    """CF.R:123-191. Inverse covariance; in the db branch `theta[[i]][3]` is the noise sd and the
    `sigma` argument is ignored unless `theta` is a plain numeric vector (CF.R:157-176)."""
    if not isinstance(x, dict):
        mat, xs = _cov_vector(x, kern, theta, sigma)
        return NamedMat(_solve(mat), xs)
    out = {}
    for i, d in split_db(x).items():
        th = theta[i] if isinstance(theta, dict) else np.r_[np.asarray(theta, dtype=np.float64), sigma]
        mat, xs = _cov_vector(d["Timestamp"], kern, th[:2], th[2])
        out[i] = NamedMat(_solve(mat), xs)
    return next(iter(out.values())) if len(out) == 1 else out


# --------------------------------------------------------------------------------------------
# CF.R:194-352 — log-likelihoods
# --------------------------------------------------------------------------------------------
def logL_GP_mod(hp, db, mean, kern, new_cov, pen_diag=None):
    """CF.R:194-216. Modified Gaussian NEGATIVE log-likelihood of one GP + trace correction."""
    hp = np.asarray(hp, dtype=np.float64)
    t = np.asarray(db["Timestamp"], dtype=np.float64)
    y = np.asarray(db["Output"], dtype=np.float64)
    mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
    if mean.shape[0] == 1:
        mean = np.repeat(mean, t.shape[0])
    sigma = hp[2] if hp.shape[0] == 3 else pen_diag
    inv = kern_to_inv(t, kern, hp[:2], sigma).m
    LL_norm = -dmvnorm(y, mean, inv, log=True)
    cor_term = 0.5 * np.sum(inv * _mat(new_cov))
    return LL_norm + cor_term


def _mat(a):
    return a.m if isinstance(a, NamedMat) else np.atleast_2d(np.asarray(a, dtype=np.float64))


def _corr_terms(mu_k_param, i, t_i):
    """the corr1/corr2 loop shared by CF.R:236-245, 304-313, 484-492, 556-565."""
    corr1 = 0.0
    corr2 = 0.0
    for k in mu_k_param["mean"]:
        tau = mu_k_param["tau_i_k"][k][i]
        mk = mu_k_param["mean"][k]
        # `filter(Timestamp %in% t_i)` keeps the grid's (sorted) order
        mean_mu_k = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], t_i)]
        corr1 = corr1 + tau * mean_mu_k
        corr2 = corr2 + tau * (np.outer(mean_mu_k, mean_mu_k) + mu_k_param["cov"][k].sub(t_i))
    return corr1, corr2


def logL_clust_multi_GP(hp, db, mu_k_param, kern):
    """CF.R:218-248. F_i(theta) for one individual."""
    hp = np.asarray(hp, dtype=np.float64)
    sigma = hp[2] if hp.shape[0] == 3 else 0.1
    t_i = np.asarray(db["Timestamp"], dtype=np.float64)
    y_i = np.asarray(db["Output"], dtype=np.float64)
    i = _key(unique_stable(db["ID"])[0])
    inv = kern_to_inv(t_i, kern, hp[:2], sigma).m
    LL_norm = -dmvnorm(y_i, np.zeros_like(y_i), inv, log=True)
    corr1, corr2 = _corr_terms(mu_k_param, i, t_i)
    return LL_norm - y_i @ inv @ corr1 + 0.5 * np.sum(inv * corr2)


def logL_GP_mod_common_hp_k(hp, db, mean, kern, new_cov, pen_diag=None):
    """CF.R:250-278. Sum over clusters of F_k with shared theta_k (db = dict of mean tibbles)."""
    hp = np.asarray(hp, dtype=np.float64)
    sigma = hp[2] if hp.shape[0] == 3 else pen_diag
    names = list(db.keys())
    t_k = np.asarray(db[names[0]]["Timestamp"], dtype=np.float64)
    inv = kern_to_inv(t_k, kern, hp[:2], sigma).m
    LL_norm = 0.0
    cor_term = 0.0
    for k in names:
        y_k = np.asarray(db[k]["Output"], dtype=np.float64)
        mk = np.atleast_1d(np.asarray(mean[k], dtype=np.float64))
        # R: rep(mean[[k]], length(t_k)) — scalar prior mean replicated
        mu = np.repeat(mk, t_k.shape[0]) if mk.shape[0] == 1 else mk
        LL_norm = LL_norm - dmvnorm(y_k, mu, inv, log=True)
        cor_term = cor_term + 0.5 * np.sum(inv * _mat(new_cov[k]))
    return LL_norm + cor_term


def logL_clust_multi_GP_common_hp_i(hp, db, mu_k_param, kern):
    """CF.R:280-325. Sum over individuals of F_i with shared theta_i."""
    hp = np.asarray(hp, dtype=np.float64)
    sigma = hp[2] if hp.shape[0] == 3 else 0.2
    sum_i = 0.0
    for i, d in split_db(db).items():
        t_i, y_i = d["Timestamp"], d["Output"]
        corr1, corr2 = _corr_terms(mu_k_param, i, t_i)
        inv = kern_to_inv(t_i, kern, hp[:2], sigma).m  # Q4: never cached (CF.R:296,315)
        LL_norm = -dmvnorm(y_i, np.zeros_like(y_i), inv, log=True)
        sum_i = sum_i + LL_norm - y_i @ inv @ corr1 + 0.5 * np.sum(inv * corr2)
    return sum_i


def logL_monitoring_VEM(hp, db, kern_i, kern_0, mu_k_param, m_k):
    """CF.R:327-352. ELBO core = -(sum_k F_k + sum_i F_i)."""
    sum_ll_k = 0.0
    for k in m_k:
        sum_ll_k += logL_GP_mod(hp["theta_k"][k], mu_k_param["mean"][k], m_k[k], kern_0,
                                mu_k_param["cov"][k])
    sum_ll_i = 0.0
    for i, d in split_db(db).items():
        sum_ll_i += logL_clust_multi_GP(hp["theta_i"][i], d, mu_k_param, kern_0)  # Q5: kern_0
    return -sum_ll_k - sum_ll_i


def BIC(hp, db, kern_0, kern_i, m_k, mu_k_param, clust=True):
    """CF.R:354-432 (clust = TRUE branch). `maotai::pdeterminant(x)$modulus` (third-party, unpinned) is taken as the
    log-determinant: the pseudo-determinant of a full-rank SPD matrix. Quirks kept: the numbers of hyper-parameters
    are counts of DISTINCT VALUES (`n_distinct(unlist(.))`, CF.R:359-361) and every cluster is centred on the FIRST
    prior mean `m_k[[1]]` (CF.R:409)."""
    if not clust:
        raise NotImplementedError("clust = FALSE needs the single-cluster MAGMA model of the sibling repository")
    list_clust = list(mu_k_param["mean"].keys())
    per_id = split_db(db)
    nb_indiv = len(per_id)
    nb_hp_i = len(np.unique(np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in hp["theta_i"].values()])))
    nb_hp_k = len(np.unique(np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in hp["theta_k"].values()])))
    nb_clust = len(np.unique(np.asarray(hp["pi_k"], dtype=np.float64)))
    pen = 0.5 * (nb_hp_i + nb_hp_k + nb_clust - 1) * np.log(nb_indiv)
    pi = np.asarray(hp["pi_k"], dtype=np.float64)
    LL_i = 0.0
    for i, d in per_id.items():
        t_i, y_i = d["Timestamp"], d["Output"]
        th = np.asarray(hp["theta_i"][i], dtype=np.float64).ravel()
        inv_i = kern_to_inv(t_i, kern_i, th[:2], th[2]).m
        sum_LL = 0.0
        LL_Z = 0.0
        for c, k in enumerate(list_clust):
            mk = mu_k_param["mean"][k]
            mu_k_ti = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], t_i)]
            tau = mu_k_param["tau_i_k"][k][i]
            cov_k_hat = mu_k_param["cov"][k].sub(t_i)
            sum_LL += tau * (dmvnorm(y_i, mu_k_ti, inv_i, log=True) - 0.5 * np.sum(inv_i * cov_k_hat))
            LL_Z += tau * (0.0 if (tau == 0 or pi[c] == 0) else np.log(pi[c] / tau))
        LL_i += sum_LL + LL_Z
    t = unique_stable(np.asarray(db["Timestamp"], dtype=np.float64))
    size_mu = t.shape[0]
    first = np.atleast_1d(np.asarray(next(iter(m_k.values())), dtype=np.float64))
    LL_k = 0.0
    for k in list_clust:
        mk = mu_k_param["mean"][k]
        mu_k_t = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], t)]
        mean_k = np.repeat(first[0], size_mu)
        thk = np.asarray(hp["theta_k"][k], dtype=np.float64)
        inv_k = kern_to_inv(t, kern_0, thk[:2], thk[2])
        cov_k_hat = mu_k_param["cov"][k].sub(inv_k.t)
        LL_k += dmvnorm(mu_k_t, mean_k, inv_k.m, log=True) - 0.5 * np.sum(inv_k.m * cov_k_hat) + \
            0.5 * (size_mu + size_mu * LOG_2PI + np.linalg.slogdet(cov_k_hat)[1])
    return LL_i + LL_k - pen


def model_selection(db, k_grid, ini_tau_by_k, ini_hp=None, kern_0=kernel_mu, kern_i=kernel, common_hp_k=True,
                    common_hp_i=True):
    """MC.R:101-135 for K >= 2 (K = 1 calls `training` of the sibling MAGMA repository, which is not in the reference).
    `ini_tau_by_k[K]` supplies the initial responsibilities (the reference's default is an unseeded k-means)."""
    res = {}
    for K in k_grid:
        if K < 2:
            raise NotImplementedError("K = 1 needs the sibling repository's training()")
        prior_mean_k = {f"K{k + 1}": 0.0 for k in range(K)}
        model = training_VEM(db, prior_mean_k, ini_hp, kern_0, kern_i, ini_tau_by_k[K], common_hp_k, common_hp_i)
        model["BIC"] = BIC(model["hp"], db, kern_0, kern_i, prior_mean_k, model["param"], True)
        res[f"K = {K}"] = model
    bics = {K: res[f"K = {K}"]["BIC"] for K in k_grid}
    res["K_max_BIC"] = max(bics, key=bics.get)
    return res


# --------------------------------------------------------------------------------------------
# CF.R:448-586 — gradients
# --------------------------------------------------------------------------------------------
def gr_GP_mod(hp, db, mean, kern, new_cov, pen_diag=None):
    """CF.R:448-475."""
    hp = np.asarray(hp, dtype=np.float64)
    y = np.asarray(db["Output"], dtype=np.float64)
    t = np.asarray(db["Timestamp"], dtype=np.float64)
    sigma = hp[2] if hp.shape[0] == 3 else pen_diag
    inv = kern_to_inv(t, kern, hp[:2], sigma).m
    prod_inv = inv @ (y - mean)
    cste_term = np.outer(prod_inv, prod_inv) + inv @ (_mat(new_cov) @ inv - np.eye(t.shape[0]))
    g_1 = 0.5 * np.trace(cste_term @ kern_to_cov(t, deriv_hp1, hp[:2], 0.0).m)
    if t.shape[0] == 1:
        g_2 = 0.0
    else:
        g_2 = 0.5 * np.trace(cste_term @ deriv_hp2(_dist2(t), hp[:2]))
    if hp.shape[0] == 3:
        g_3 = hp[2] * np.trace(cste_term)
        return -np.array([g_1, g_2, g_3])
    return -np.array([g_1, g_2])


def gr_clust_multi_GP(hp, db, mu_k_param, kern):
    """CF.R:477-508."""
    hp = np.asarray(hp, dtype=np.float64)
    t_i = np.asarray(db["Timestamp"], dtype=np.float64)
    y_i = np.asarray(db["Output"], dtype=np.float64)
    i = _key(unique_stable(db["ID"])[0])
    corr1, corr2 = _corr_terms(mu_k_param, i, t_i)
    sigma = hp[2] if hp.shape[0] == 3 else 0.1
    inv = kern_to_inv(t_i, kern, hp[:2], sigma).m
    prod_inv = inv @ y_i
    cste_term = np.outer(prod_inv - 2.0 * inv @ corr1, prod_inv) + inv @ (corr2 @ inv - np.eye(t_i.shape[0]))
    g_1 = 0.5 * np.trace(cste_term @ kern_to_cov(t_i, deriv_hp1, hp[:2], 0.0).m)
    g_2 = 0.5 * np.trace(cste_term @ deriv_hp2(_dist2(t_i), hp[:2]))
    if hp.shape[0] == 3:
        g_3 = hp[2] * np.trace(cste_term)
        return -np.array([g_1, g_2, g_3])
    return -np.array([g_1, g_2])


def gr_GP_mod_common_hp_k(hp, db, mean, kern, new_cov, pen_diag=None):
    """CF.R:510-537."""
    hp = np.asarray(hp, dtype=np.float64)
    names = list(db.keys())
    t_k = np.asarray(db[names[0]]["Timestamp"], dtype=np.float64)
    sigma = hp[2] if hp.shape[0] == 3 else pen_diag
    inv = kern_to_inv(t_k, kern, hp[:2], sigma).m
    d1 = kern_to_cov(t_k, deriv_hp1, hp[:2], 0.0).m
    d2 = deriv_hp2(_dist2(t_k), hp[:2])
    g_1 = g_2 = g_3 = 0.0
    for k in names:
        y_k = np.asarray(db[k]["Output"], dtype=np.float64)
        prod_inv = inv @ (y_k - mean[k])
        cste_term = np.outer(prod_inv, prod_inv) + inv @ (_mat(new_cov[k]) @ inv - np.eye(t_k.shape[0]))
        g_1 += 0.5 * np.trace(cste_term @ d1)
        g_2 += 0.5 * np.trace(cste_term @ d2)
        if hp.shape[0] == 3:
            g_3 += hp[2] * np.trace(cste_term)
    if hp.shape[0] == 3:
        return -np.array([g_1, g_2, g_3])
    return -np.array([g_1, g_2])


def gr_clust_multi_GP_common_hp_i(hp, db, mu_k_param, kern):
    """CF.R:539-586."""
    hp = np.asarray(hp, dtype=np.float64)
    sigma = hp[2] if hp.shape[0] == 3 else 0.2
    g_1 = g_2 = g_3 = 0.0
    for i, d in split_db(db).items():
        t_i, y_i = d["Timestamp"], d["Output"]
        corr1, corr2 = _corr_terms(mu_k_param, i, t_i)
        inv = kern_to_inv(t_i, kern, hp[:2], sigma).m
        prod_inv = inv @ y_i
        cste_term = np.outer(prod_inv - 2.0 * inv @ corr1, prod_inv) + inv @ (corr2 @ inv - np.eye(t_i.shape[0]))
        g_1 += 0.5 * np.trace(cste_term @ kern_to_cov(t_i, deriv_hp1, hp[:2], 0.0).m)
        g_2 += 0.5 * np.trace(cste_term @ deriv_hp2(_dist2(t_i), hp[:2]))
        if hp.shape[0] == 3:
            g_3 += hp[2] * np.trace(cste_term)
    if hp.shape[0] == 3:
        return -np.array([g_1, g_2, g_3])
    return -np.array([g_1, g_2])


# --------------------------------------------------------------------------------------------
# CF.R:589-786 — E step, M step, update helpers
# --------------------------------------------------------------------------------------------
def update_inv_VEM(prior_inv, list_inv_i, tau_i_k):
    """CF.R:690-707. A_k = K_k^-1 + sum_i tau_ik P_i' Psi_i^-1 P_i (name-indexed scatter)."""
    new_inv = NamedMat(prior_inv.m.copy(), prior_inv.t)
    for x, inv_i in list_inv_i.items():
        common = inv_i.t[np.isin(inv_i.t, new_inv.t)]  # intersect() keeps inv_i's order
        r = new_inv.pos(common)
        new_inv.m[np.ix_(r, r)] += tau_i_k[x] * inv_i.sub(common)
    return new_inv


def update_mean_VEM(prior_mean, prior_inv, list_inv_i, list_value_i, tau_i_k):
    """CF.R:709-731. w_k = K_k^-1 m_k + sum_i tau_ik P_i' Psi_i^-1 y_i."""
    prior_mean = np.atleast_1d(np.asarray(prior_mean, dtype=np.float64))
    if prior_mean.shape[0] == 1:
        prior_mean = np.repeat(prior_mean, prior_inv.m.shape[1])
    weighted_mean = prior_inv.m @ prior_mean
    for i, inv_i in list_inv_i.items():
        weighted_i = tau_i_k[i] * (inv_i.m @ list_value_i[i])
        keep = np.isin(inv_i.t, prior_inv.t)
        r = prior_inv.pos(inv_i.t[keep])
        weighted_mean[r] += weighted_i[keep]
    return weighted_mean


def update_tau_i_k_VEM(db, m_k, mean_k, cov_k, kern_i, hp, pi_k):
    """CF.R:733-762. Responsibilities with the log-sum-exp trick (max over logL only, pi after)."""
    per_id = split_db(db)
    names_k = list(m_k.keys())
    mat_logL = np.full((len(names_k), len(per_id)), np.nan)
    for c_i, (i, d) in enumerate(per_id.items()):
        t_i = d["Timestamp"]
        for c_k, k in enumerate(names_k):
            mk = mean_k[k]
            mean_sel = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], t_i)]
            mat_logL[c_k, c_i] = -logL_GP_mod(hp["theta_i"][i], d, mean_sel, kern_i, cov_k[k].sub(t_i))
    mat_L = np.exp(mat_logL - mat_logL.max(axis=0, keepdims=True))
    w = np.asarray(pi_k, dtype=np.float64)[:, None] * mat_L  # positional pi_k (Q3)
    w = w / w.sum(axis=0, keepdims=True)
    ids = list(per_id.keys())
    return {k: {i: w[c_k, c_i] for c_i, i in enumerate(ids)} for c_k, k in enumerate(names_k)}, mat_logL


def e_step_VEM(db, m_k, kern_0, kern_i, hp, old_tau_i_k, return_logL=False):
    """CF.R:589-629. Hyper-posteriors of the K mean processes, then responsibilities."""
    pi_k = hp["pi_k"]
    all_t = np.sort(unique_stable(np.asarray(db["Timestamp"], dtype=np.float64)))
    names_k = list(m_k.keys())
    # t_clust tibble: every cluster gets the full grid; theta_k[[k]][3] is the jitter sd
    inv_k = {k: kern_to_inv(all_t, kern_0, np.asarray(hp["theta_k"][k])[:2], np.asarray(hp["theta_k"][k])[2])
             for k in names_k}
    inv_i = kern_to_inv(db, kern_i, hp["theta_i"], 0)
    if isinstance(inv_i, NamedMat):
        inv_i = {_key(unique_stable(db["ID"])[0]): inv_i}
    value_i = {i: d["Output"] for i, d in split_db(db).items()}
    cov_k = {}
    for k in names_k:
        new_inv = update_inv_VEM(inv_k[k], inv_i, old_tau_i_k[k])
        cov_k[k] = NamedMat(_solve(new_inv.m), all_t)
    mean_k = {}
    for k in names_k:
        weighted_mean = update_mean_VEM(m_k[k], inv_k[k], inv_i, value_i, old_tau_i_k[k])
        mean_k[k] = {"Timestamp": all_t, "Output": cov_k[k].m @ weighted_mean}
    tau_i_k, mat_logL = update_tau_i_k_VEM(db, m_k, mean_k, cov_k, kern_i, hp, pi_k)
    out = {"mean": mean_k, "cov": cov_k, "tau_i_k": tau_i_k}
    if return_logL:
        out["mat_logL"] = mat_logL
    return out


def _lbfgsb(fn, gr, par, args):
    """stand-in for optimr::opm(..., method='L-BFGS-B', control=list(kkt=F)) (CF.R:647,656,666,675).

    R's optim defaults: lmm=5, factr=1e7, pgtol=0, maxit=100. scipy ships a different L-BFGS-B
    build, so trajectories are NOT comparable with R ("parity unpinned" for this boundary)."""
    from scipy.optimize import minimize

    res = minimize(fn, np.asarray(par, dtype=np.float64), args=args, jac=gr, method="L-BFGS-B",
                   options={"maxcor": 5, "ftol": 1e7 * np.finfo(float).eps, "gtol": 0.0, "maxiter": 100})
    return res.x, res


def m_step_VEM(db, old_hp, list_mu_param, kern_0, kern_i, m_k, common_hp_k, common_hp_i, stats=None):
    """CF.R:631-687. New theta_i, theta_k (L-BFGS-B) and pi_k."""
    list_ID_k = list(m_k.keys())
    per_id = split_db(db)
    list_ID_i = list(per_id.keys())
    nfev = 0
    t1 = time.perf_counter()
    if common_hp_i:
        par, res = _lbfgsb(logL_clust_multi_GP_common_hp_i, gr_clust_multi_GP_common_hp_i,
                           old_hp["theta_i"][list_ID_i[0]], (db, list_mu_param, kern_i))
        nfev += res.nfev
        new_theta_i = {i: par.copy() for i in list_ID_i}
    else:
        new_theta_i = {}
        for i in list_ID_i:
            par, res = _lbfgsb(logL_clust_multi_GP, gr_clust_multi_GP, old_hp["theta_i"][i],
                               (per_id[i], list_mu_param, kern_i))
            nfev += res.nfev
            new_theta_i[i] = par
    t2 = time.perf_counter()
    pen_diag = float(np.mean([new_theta_i[i][2] for i in list_ID_i]))
    nfev_k = 0
    if common_hp_k:
        par, res = _lbfgsb(logL_GP_mod_common_hp_k, gr_GP_mod_common_hp_k,
                           np.asarray(old_hp["theta_k"][list_ID_k[0]], dtype=np.float64),
                           (list_mu_param["mean"], m_k, kern_0, list_mu_param["cov"], pen_diag))
        # NB (reference behaviour): par has 3 entries here, so sigma = par[3] is optimised and
        # then DISCARDED: opm(...)[1,1:2] keeps (a, b) and pen_diag replaces the third (CF.R:666-668)
        nfev_k += res.nfev
        new_theta_k = {k: np.r_[par[:2], pen_diag] for k in list_ID_k}
    else:
        new_theta_k = {}
        for k in list_ID_k:
            par, res = _lbfgsb(logL_GP_mod, gr_GP_mod, np.asarray(old_hp["theta_k"][k], dtype=np.float64)[:2],
                               (list_mu_param["mean"][k], m_k[k], kern_0, list_mu_param["cov"][k], pen_diag))
            nfev_k += res.nfev
            new_theta_k[k] = np.r_[par[:2], pen_diag]
    pi_k = np.array([np.mean(list(list_mu_param["tau_i_k"][k].values())) for k in list_mu_param["tau_i_k"]])
    if stats is not None:
        stats["nfev_i"] = stats.get("nfev_i", 0) + nfev
        stats["nfev_k"] = stats.get("nfev_k", 0) + nfev_k
        stats["t_i"] = stats.get("t_i", 0.0) + (t2 - t1)                      # 'indiv:' of CF.R:685
        stats["t_k"] = stats.get("t_k", 0.0) + (time.perf_counter() - t2)     # 'mu_k:'  of CF.R:685
    return {"theta_k": new_theta_k, "theta_i": new_theta_i, "pi_k": pi_k}


def update_tau_star_k_EM(db, mean_k, cov_k, kern, hp, pi_k):
    """CF.R:764-786. Responsibilities of a new individual."""
    names_k = list(mean_k.keys())
    t_i = np.asarray(db["Timestamp"], dtype=np.float64)
    y = np.asarray(db["Output"], dtype=np.float64)
    pi = np.array([pi_k[k] for k in names_k], dtype=np.float64) if isinstance(pi_k, dict) else np.asarray(pi_k)
    hp = np.asarray(hp, dtype=np.float64)
    mat_logL = np.full(len(names_k), np.nan)
    for c_k, k in enumerate(names_k):
        mk = mean_k[k]
        mean = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], t_i)]
        cov = kern_to_cov(t_i, kern, hp[:2], hp[2]).m + cov_k[k].sub(t_i)
        inv = _solve(cov)
        mat_logL[c_k] = dmvnorm(y, mean, inv, log=True)
    mat_L = np.exp(mat_logL - mat_logL.max())
    w = pi * mat_L / np.sum(pi * mat_L)
    return {k: w[c_k] for c_k, k in enumerate(names_k)}


# --------------------------------------------------------------------------------------------
# MC.R — drivers
# --------------------------------------------------------------------------------------------
def _logdet_cov(x, mirror_underflow):
    """MC.R:53 computes log(det(x)), which underflows to -Inf for N >~ 210 (quirk Q8)."""
    if mirror_underflow:
        with np.errstate(divide="ignore"):
            return float(np.log(np.linalg.det(x)))
    return float(np.linalg.slogdet(x)[1])


def elbo_VEM(hp, db, kern_i, kern_0, param, m_k, mirror_underflow=False):
    """MC.R:51-53. logL_monitoring_VEM + 0.5*(K*N + sum_k log det C_k)."""
    covs = list(param["cov"].values())
    return logL_monitoring_VEM(hp, db, kern_i, kern_0, param, m_k) + 0.5 * (
        len(covs) * covs[0].m.shape[0] + sum(_logdet_cov(c.m, mirror_underflow) for c in covs))


def training_VEM(db, prior_mean_k, ini_hp=None, kern_0=kernel_mu, kern_i=kernel, ini_tau_i_k=None,
                 common_hp_k=True, common_hp_i=True, n_loop_max=15, verbose=False,
                 mirror_underflow=False, stats=None):
    """MC.R:16-99. VEM loop with the reference's stopping rule (eps < 1e-2) and return shape.

    `ini_tau_i_k` must be supplied: the reference's default is an unseeded `stats::kmeans`
    (MC.R:37), which is outside the pinned path."""
    if ini_hp is None:
        ini_hp = {"theta_k": np.array([1.0, 1.0, 0.2]), "theta_i": np.array([1.0, 1.0, 0.2])}
    if ini_tau_i_k is None:
        raise ValueError("ini_tau_i_k is required (k-means initialisation is not part of the oracle)")
    list_ID = list(split_db(db).keys())
    ID_k = list(prior_mean_k.keys())
    hp = {"theta_k": {k: np.asarray(ini_hp["theta_k"], dtype=np.float64) for k in ID_k},
          "theta_i": {i: np.asarray(ini_hp["theta_i"], dtype=np.float64) for i in list_ID}}
    cv = False
    tau_i_k = ini_tau_i_k
    hp["pi_k"] = np.array([np.mean(list(tau_i_k[k].values())) for k in tau_i_k])
    logLL_monitoring = -np.inf
    t1 = time.time()
    seq_logL = []
    new_hp = hp
    param = None
    for it in range(n_loop_max):
        param = e_step_VEM(db, prior_mean_k, kern_0, kern_i, hp, tau_i_k)
        new_logLL_monitoring = elbo_VEM(hp, db, kern_i, kern_0, param, prior_mean_k, mirror_underflow)
        if verbose:
            print(it + 1, new_logLL_monitoring)
        new_hp = m_step_VEM(db, hp, param, kern_0, kern_i, prior_mean_k, common_hp_k, common_hp_i, stats)
        if _any_nan(new_hp):
            break
        logL_new = logL_monitoring_VEM(new_hp, db, kern_i, kern_0, param, prior_mean_k)
        eps = (logL_new - logL_monitoring_VEM(hp, db, kern_i, kern_0, param, prior_mean_k)) / abs(logL_new)
        if verbose:
            print("eps", eps)
        seq_logL.append(new_logLL_monitoring)
        if eps < 1e-2:
            if eps > 0:
                hp = new_hp
            cv = True
            break
        hp = new_hp
        tau_i_k = param["tau_i_k"]
        logLL_monitoring = new_logLL_monitoring
    t2 = time.time()
    return {"hp": new_hp, "convergence": cv, "param": param, "Time_train": t2 - t1,
            "logL_iterations": np.array(seq_logL), "n_iter": len(seq_logL)}


def _any_nan(hp):
    vals = [np.asarray(v, dtype=np.float64) for v in hp["theta_k"].values()]
    vals += [np.asarray(v, dtype=np.float64) for v in hp["theta_i"].values()]
    vals.append(np.asarray(hp["pi_k"], dtype=np.float64))
    return any(np.isnan(v).any() for v in vals)


def posterior_mu_k(db, timestamps, m_k, kern_0, kern_i, list_hp):
    """MC.R:196-233. Hyper-posterior on the prediction grid; theta_k[3] forced to 0.1, tau frozen."""
    hp_i = list_hp["hp"]["theta_i"]
    hp_k = {k: np.r_[np.asarray(v, dtype=np.float64)[:2], 0.1] for k, v in list_hp["hp"]["theta_k"].items()}
    tau_i_k = list_hp["param"]["tau_i_k"]
    timestamps = np.asarray(timestamps, dtype=np.float64)
    inv_k = {}
    for k in hp_k:
        # kern_to_inv sorts the grid (CF.R:166); `timestamps` is expected sorted (MC.R:323)
        inv_k[k] = kern_to_inv(timestamps, kern_0, hp_k[k][:2], hp_k[k][2])
    inv_i = kern_to_inv(db, kern_i, hp_i, 0)
    if isinstance(inv_i, NamedMat):
        inv_i = {_key(unique_stable(db["ID"])[0]): inv_i}
    value_i = {i: d["Output"] for i, d in split_db(db).items()}
    cov_k = {}
    for k in hp_k:
        new_inv = update_inv_VEM(inv_k[k], inv_i, tau_i_k[k])
        cov_k[k] = NamedMat(_solve(new_inv.m), inv_k[k].t)
    mean_k = {}
    for k in hp_k:
        weighted_mean = update_mean_VEM(m_k[k], inv_k[k], inv_i, value_i, tau_i_k[k])
        mean_k[k] = {"Timestamp": timestamps, "Output": cov_k[k].m @ weighted_mean}
    return {"mean": mean_k, "cov": cov_k, "tau_i_k": tau_i_k}


def new_indiv_logl(db, mean_k, cov_k, kern, hp):
    """the K log-densities log N(y*; m_k(t*), Psi_hp(t*) + C_k(t*, t*)) that both steps of train_new_gp_EM are made of
    (CF.R:776-780 for the E step, MC.R:156-161 for the M step); NaN where the covariance is not positive definite"""
    names_k = list(mean_k.keys())
    t_i = np.asarray(db["Timestamp"], dtype=np.float64)
    y = np.asarray(db["Output"], dtype=np.float64)
    hp = np.asarray(hp, dtype=np.float64)
    out = np.full(len(names_k), np.nan)
    for c_k, k in enumerate(names_k):
        mk = mean_k[k]
        mean = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], t_i)]
        cov = kern_to_cov(t_i, kern, hp[:2], hp[2]).m + cov_k[k].sub(t_i)
        try:
            np.linalg.cholesky(cov)
        except np.linalg.LinAlgError:
            continue
        out[c_k] = dmvnorm(y, mean, np.linalg.inv(cov), log=True)
    return out


NEW_INDIV_NDEPS = 1e-3   # stats::optim's default finite-difference step (no `gr` is passed at MC.R:166)


def new_indiv_objective(hp, tau, db, mean_k, cov_k, kern):
    """REPAIRED objective of the M step of train_new_gp_EM (MC.R:151-165): - sum_k tau_k log N(y*; m_k, Psi_hp + C_k).
    Upstream the function body reads an undefined `t` and adds `db$Timestamp` to every term (Q9); what is restated here
    is the expected complete-data log-likelihood the surrounding EM needs. A cluster of weight exactly 0 does not enter."""
    ll = new_indiv_logl(db, mean_k, cov_k, kern, hp)
    tau = np.asarray(tau, dtype=np.float64)
    return float(-np.sum(tau[tau != 0.0] * ll[tau != 0.0]))


def new_indiv_gradient(hp, tau, db, mean_k, cov_k, kern):
    """optim's numerical gradient: central differences with ndeps = 1e-3"""
    hp = np.asarray(hp, dtype=np.float64)
    g = np.zeros(3)
    for j in range(3):
        e = np.zeros(3)
        e[j] = NEW_INDIV_NDEPS
        g[j] = (new_indiv_objective(hp + e, tau, db, mean_k, cov_k, kern)
                - new_indiv_objective(hp - e, tau, db, mean_k, cov_k, kern)) / (2.0 * NEW_INDIV_NDEPS)
    return g


def train_new_gp_EM(db, param_mu_k, ini_hp=None, kern_i=kernel, hp_i=None, n_loop_max=25, return_loops=False):
    """MC.R:137-192. Fixed-hp branch as upstream; the free-hp branch (hp_i = None) with the repaired objective above
    ("parity unpinned": upstream cannot run it, and the optimiser is scipy's L-BFGS-B as everywhere in this oracle)."""
    pi_k = {k: float(np.mean(list(v.values()))) for k, v in param_mu_k["tau_i_k"].items()}
    if hp_i is None:
        mean_k, cov_k = param_mu_k["mean"], param_mu_k["cov"]
        if ini_hp is None:
            ini_hp = {"theta_i": np.array([1.0, 1.0, 0.2])}
        hp = np.asarray(ini_hp["theta_i"], dtype=np.float64)[:3].copy()
        new_hp, tau_k, loops = hp, None, 0
        for _ in range(n_loop_max):
            tau_k = update_tau_star_k_EM(db, mean_k, cov_k, kern_i, hp, pi_k)            # MC.R:149
            tau = np.array([tau_k[k] for k in mean_k])
            new_hp, _ = _lbfgsb(new_indiv_objective, new_indiv_gradient, hp, (tau, db, mean_k, cov_k, kern_i))   # MC.R:166
            loops += 1
            eps = float(np.sum(np.abs(new_hp - hp)))                                       # MC.R:175
            if 0.0 < eps < 1e-3:
                break
            hp = new_hp
        out = {"theta_new": new_hp, "tau_k": tau_k}
        if return_loops:
            out["loops"] = loops
        return out
    new_hp = np.asarray(next(iter(hp_i["hp"]["theta_i"].values())), dtype=np.float64)
    tau_k = update_tau_star_k_EM(db, param_mu_k["mean"], param_mu_k["cov"], kern_i, new_hp, pi_k)
    return {"theta_new": new_hp, "tau_k": tau_k}


def pred_gp_clust(db, timestamps, list_mu, kern, hp):
    """MC.R:235-274. Per-cluster GP predictive mean/variance at `timestamps`."""
    tn = np.asarray(db["Timestamp"], dtype=np.float64)
    yn = np.asarray(db["Output"], dtype=np.float64)
    if timestamps is None:
        timestamps = np.linspace(tn.min(), tn.max(), 500)
    timestamps = np.asarray(timestamps, dtype=np.float64)
    theta = np.asarray(hp["theta_new"], dtype=np.float64)[:2]
    sigma = float(np.asarray(hp["theta_new"], dtype=np.float64)[2])
    pred = {}
    for k in list_mu["mean"]:
        mk = list_mu["mean"][k]
        mean_mu_obs = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], tn)]
        mean_mu_pred = np.asarray(mk["Output"])[np.isin(mk["Timestamp"], timestamps)]
        cov_mu = list_mu["cov"][k]
        cov_tn_tn = kern_to_cov(tn, kern, theta, sigma).m + cov_mu.sub(tn)
        inv_mat = _solve(cov_tn_tn)
        cov_tn_t = kern(mat_dist(tn, timestamps), theta) + cov_mu.sub(tn, timestamps)
        cov_t_t = kern_to_cov(timestamps, kern, theta, sigma).m + cov_mu.sub(timestamps)
        pred[k] = {
            "Timestamp": timestamps,
            "Mean": mean_mu_pred + cov_tn_t.T @ inv_mat @ (yn - mean_mu_obs),
            "Var": np.diag(cov_t_t - cov_tn_t.T @ inv_mat @ cov_tn_t).copy(),
            "tau_k": hp["tau_k"][k],
        }
    return pred


def full_algo_clust(db, new_db, timestamps, kern_i, ini_tau_i_k=None, common_hp_k=True, common_hp_i=True,
                    prior_mean=None, kern_0=None, list_hp=None, mu_k=None, ini_hp=None, hp_new_i=None):
    """MC.R:301-347."""
    if list_hp is None:
        list_hp = training_VEM(db, prior_mean, ini_hp, kern_0, kern_i, ini_tau_i_k, common_hp_k, common_hp_i)
    t_pred = np.sort(unique_stable(np.r_[np.asarray(timestamps, dtype=np.float64),
                                         unique_stable(db["Timestamp"]), unique_stable(new_db["Timestamp"])]))
    if mu_k is None:
        mu_k = posterior_mu_k(db, t_pred, prior_mean, kern_0, kern_i, list_hp)
    if hp_new_i is None:
        hp_new_i = train_new_gp_EM(new_db, mu_k, ini_hp, kern_i, hp_i=list_hp if common_hp_i else None)
    pred = pred_gp_clust(new_db, timestamps, mu_k, kern_i, hp_new_i)
    return {"Prediction": pred, "Mean_processes": mu_k,
            "Hyperparameters": {"other_i": list_hp, "new_i": hp_new_i},
            "logL_iterations": list_hp.get("logL_iterations"),
            "ari_iterations": list_hp.get("ARI_iterations")}   # MC.R:345: never set by training_VEM, i.e. NULL


def pred_max_cluster(tau_i_k):
    """MC.R:349-355. Hard assignment: which.max per individual (first maximum wins), 1-based."""
    names_k = list(tau_i_k.keys())
    ids = list(tau_i_k[names_k[0]].keys())
    mat = np.array([[tau_i_k[k][i] for k in names_k] for i in ids])
    return np.argmax(mat, axis=1) + 1


# --------------------------------------------------------------------------------------------
# conversions used by tests/bench to feed the same inputs to the CUDA path
# --------------------------------------------------------------------------------------------
def tau_to_array(tau_i_k):
    """dict[k][i] -> (M, K) array (row i = individual, col k = cluster)."""
    names_k = list(tau_i_k.keys())
    ids = list(tau_i_k[names_k[0]].keys())
    return np.array([[tau_i_k[k][i] for k in names_k] for i in ids], dtype=np.float64)


def tau_from_array(arr, ids, names_k):
    return {k: {i: float(arr[r, c]) for r, i in enumerate(ids)} for c, k in enumerate(names_k)}
