## magmaclust_b200.R — drop-in definitions of the reference's hot-path functions on top of the .Call shim.
##
## Usage in the reference:
##   source('Computing_functions.R'); source('MAGMAclust.R'); source('r/R/magmaclust_b200.R')
## The definitions below mask the R ones; signatures and return shapes are those of Computing_functions.R (CF.R) and
## MAGMAclust.R (MC.R). R is not installed in the build image: this file has never been run by R itself. What is executed
## in CI is the C side it calls (r/src/mcb_shim.c against tests/r_stub, tests/test_r_shim.py: every .Call below names a shim
## entry point of the same arity) and the same C ABI through ctypes (tests/).
##
## Resident state. R passes values, the library keeps device-resident state. Results that describe resident state carry an
## attribute ('mcb_token' on `param`, 'mcb_pred_token' on `list_mu`); a wrapper that receives an object without the current
## token uploads it (mcbR_set_posterior / mcbR_set_pred_posterior) instead of trusting what is resident. `db` is recognised
## by a content hash of its three columns (mcbR_hash).

.mcb <- new.env()

mcb_session <- function(device = 0L, shim = 'r/src/mcb_shim.so') {
  if (is.null(.mcb$h)) {
    dyn.load(shim)
    .mcb$h <- .Call('mcbR_create', as.integer(device))
    .mcb$epoch <- 0L
  }
  .mcb$h
}

.check_kernels <- function(...) for (k in list(...))
  if (!identical(k, kernel) && !identical(k, kernel_mu))
    stop('magmaclust_b200: only the squared-exponential kernels kernel / kernel_mu are implemented on the GPU')

## ---- data -----------------------------------------------------------------------------------------------------------------
## CSR canonicalisation: replaces db %>% filter(ID == i) and the 'X<t>' dimnames (CF.R:83,146,601)
digest_db <- function(db) paste(.Call('mcbR_hash', db$ID), .Call('mcbR_hash', as.numeric(db$Timestamp)),
                                .Call('mcbR_hash', as.numeric(db$Output)))
mcb_set_db <- function(db) {
  h <- mcb_session()
  key <- digest_db(db)
  if (identical(.mcb$db_key, key)) return(invisible(.mcb$canon))
  ids <- unique(db$ID)
  all_t <- sort(unique(db$Timestamp))
  ord <- order(match(db$ID, ids))                         # blocks in unique(db$ID) order, rows kept (stable)
  counts <- as.integer(table(factor(db$ID, levels = ids)))
  canon <- list(ids = ids, all_t = all_t, offsets = c(0L, as.integer(cumsum(counts))),
                grid_idx = as.integer(match(db$Timestamp[ord], all_t) - 1L), y = as.numeric(db$Output[ord]))
  .Call('mcbR_set_data', h, canon$offsets, canon$grid_idx, canon$y, as.numeric(all_t))
  .mcb$db_key <- key; .mcb$canon <- canon; .mcb$token <- NULL
  invisible(canon)
}

## ---- packing (R lists <-> the flat arrays of include/mcb.h) ------------------------------------------------------------------
.pack_theta_k <- function(hp, names_k) sapply(names_k, function(k) c(as.numeric(unlist(hp$theta_k[[k]])), 0)[1:3])   # 3 x K
.pack_theta_i <- function(hp, ids) sapply(as.character(ids), function(i) as.numeric(unlist(hp$theta_i[[i]]))[1:3])     # 3 x M
.pack_tau <- function(tau_i_k, names_k, ids)                                                                        # K x M
  matrix(sapply(as.character(ids), function(i) sapply(names_k, function(k) tau_i_k[[k]][[i]])), nrow = length(names_k))
.unpack_tau <- function(tau, names_k, ids)
  setNames(lapply(seq_along(names_k), function(k) setNames(as.list(tau[k, ]), ids)), names_k)
.pack_mk <- function(m_k, names_k, N) {
  v <- lapply(names_k, function(k) as.numeric(m_k[[k]]))
  if (all(sapply(v, length) == 1)) unlist(v) else unlist(lapply(v, function(x) if (length(x) == 1) rep(x, N) else x))
}
.set_state <- function(theta_k, theta_i, pi_k, mk, tau) {
  .Call('mcbR_set_state', mcb_session(), theta_k, theta_i, as.numeric(pi_k), as.numeric(mk), tau)
  .mcb$theta_i <- theta_i; .mcb$theta_k <- theta_k; .mcb$mk <- as.numeric(mk); .mcb$token <- NULL
}
.new_token <- function() { .mcb$epoch <- .mcb$epoch + 1L; .mcb$epoch }

## list(mean, cov, tau_i_k) of CF.R:628 from the arrays the shim returns (mean N x K, cov N x N x K, tau K x M)
.param_out <- function(mean, cov, tau, names_k, cn) {
  nm <- paste0('X', cn$all_t)
  mean_k <- lapply(seq_along(names_k), function(k) tibble::tibble(Timestamp = cn$all_t, Output = mean[, k]))
  cov_k <- lapply(seq_along(names_k), function(k) { m <- cov[, , k]; dimnames(m) <- list(nm, nm); m })
  out <- list(mean = setNames(mean_k, names_k), cov = setNames(cov_k, names_k), tau_i_k = .unpack_tau(tau, names_k, cn$ids))
  .mcb$token <- .new_token()
  attr(out, 'mcb_token') <- .mcb$token
  out
}

## make (hp, m_k, param) the resident state; no copies when `param` is what the last E-step of this session left there
.ensure_posterior <- function(hp, m_k, param, cn) {
  names_k <- names(param$mean)
  theta_k <- .pack_theta_k(hp, names_k); theta_i <- .pack_theta_i(hp, cn$ids)
  tok <- attr(param, 'mcb_token')
  if (!is.null(tok) && identical(tok, .mcb$token) && identical(theta_k, .mcb$theta_k) && identical(theta_i, .mcb$theta_i))
    return(names_k)
  tau <- .pack_tau(param$tau_i_k, names_k, cn$ids)
  pi_k <- if (is.null(hp$pi_k)) rep(1 / length(names_k), length(names_k)) else hp$pi_k
  .set_state(theta_k, theta_i, pi_k, .pack_mk(m_k, names_k, length(cn$all_t)), tau)
  mean <- sapply(names_k, function(k) as.numeric(param$mean[[k]]$Output))                    # N x K
  cov <- array(unlist(lapply(names_k, function(k) as.numeric(param$cov[[k]]))), c(nrow(mean), nrow(mean), length(names_k)))
  .Call('mcbR_set_posterior', mcb_session(), mean, cov, tau)
  .mcb$token <- NULL
  names_k
}

## ---- building blocks (CF.R:10-191) ---------------------------------------------------------------------------------------------
kern_to_cov <- function(x, kern = kernel, theta = c(1, 0.5), sigma = 0.2) {
  .check_kernels(kern)
  if (is.vector(x)) {
    x <- sort(x)
    m <- .Call('mcbR_kern', mcb_session(), as.numeric(x), as.numeric(theta)[1:2], as.numeric(sigma), FALSE)
    dimnames(m) <- list(paste0('X', x), paste0('X', x)); m
  } else {
    floop <- function(i) { th <- if (is.list(theta)) unlist(theta[[i]]) else c(theta, sigma)
      kern_to_cov(x$Timestamp[x$ID == i], kern, th[1:2], th[[3]]) }
    res <- sapply(unique(x$ID), floop, simplify = FALSE, USE.NAMES = TRUE)
    if (length(res) == 1) res[[1]] else res
  }
}
kern_to_inv <- function(x, kern = kernel, theta = c(1, 0.5), sigma = 0.2) {
  .check_kernels(kern)
  if (is.vector(x)) {
    x <- sort(x)
    m <- .Call('mcbR_kern', mcb_session(), as.numeric(x), as.numeric(theta)[1:2], as.numeric(sigma), TRUE)
    attr(m, 'logdet') <- NULL
    dimnames(m) <- list(paste0('X', x), paste0('X', x)); m
  } else {
    floop <- function(i) { th <- if (is.list(theta)) unlist(theta[[i]]) else c(theta, sigma)
      kern_to_inv(x$Timestamp[x$ID == i], kern, th[1:2], th[[3]]) }
    res <- sapply(unique(x$ID), floop, simplify = FALSE, USE.NAMES = TRUE)
    if (length(res) == 1) res[[1]] else res
  }
}
dmvnorm <- function(x, mu, inv_Sigma, log = FALSE, tol = 1e-06)   # tol: unused by the reference as well (CF.R:10)
  .Call('mcbR_dmvnorm', mcb_session(), as.numeric(x), as.numeric(mu), as.numeric(inv_Sigma), log)

## ---- E step / M step (CF.R:589-687, 733-762) ------------------------------------------------------------------------------------
e_step_VEM <- function(db, m_k, kern_0, kern_i, hp, old_tau_i_k) {
  .check_kernels(kern_0, kern_i)
  cn <- mcb_set_db(db); names_k <- names(m_k)
  theta_k <- .pack_theta_k(hp, names_k); theta_i <- .pack_theta_i(hp, cn$ids)
  tau <- .pack_tau(old_tau_i_k, names_k, cn$ids)
  mk <- .pack_mk(m_k, names_k, length(cn$all_t))
  res <- .Call('mcbR_e_step_VEM', mcb_session(), theta_k, theta_i, as.numeric(hp$pi_k), as.numeric(mk), tau, length(cn$all_t))
  .mcb$theta_i <- theta_i; .mcb$theta_k <- theta_k; .mcb$mk <- as.numeric(mk)
  .param_out(res[[1]], res[[2]], res[[3]], names_k, cn)
}

update_tau_i_k_VEM <- function(db, m_k, mean_k, cov_k, kern_i, hp, pi_k) {
  .check_kernels(kern_i)
  cn <- mcb_set_db(db); names_k <- names(m_k); K <- length(names_k); M <- length(cn$ids)
  theta_k <- if (is.null(hp$theta_k)) matrix(c(1, 1, 0.2), 3, K) else .pack_theta_k(hp, names_k)     # unused by the update
  .set_state(theta_k, .pack_theta_i(hp, cn$ids), pi_k, .pack_mk(m_k, names_k, length(cn$all_t)), matrix(1 / K, K, M))
  mean <- sapply(names_k, function(k) as.numeric(mean_k[[k]]$Output))
  cov <- array(unlist(lapply(names_k, function(k) as.numeric(cov_k[[k]]))), c(nrow(mean), nrow(mean), K))
  tau <- .Call('mcbR_update_tau', mcb_session(), mean, cov, K, M)
  .unpack_tau(tau, names_k, cn$ids)
}

## m_step_VEM with the library's own batched L-BFGS (one kernel for all per-individual optimisations, CF.R:653-659)
m_step_VEM <- function(db, old_hp, list_mu_param, kern_0, kern_i, m_k, common_hp_k, common_hp_i) {
  .check_kernels(kern_0, kern_i)
  cn <- mcb_set_db(db)
  names_k <- .ensure_posterior(old_hp, m_k, list_mu_param, cn)
  res <- .Call('mcbR_m_step', mcb_session(), common_hp_k, common_hp_i, length(names_k), length(cn$ids))
  theta_k <- lapply(seq_along(names_k), function(k) res[[1]][, k])
  theta_i <- lapply(seq_along(cn$ids), function(i) res[[2]][, i])
  list(theta_k = setNames(theta_k, names_k), theta_i = setNames(theta_i, cn$ids), pi_k = setNames(res[[3]], names_k))
}

## ---- optimiser callbacks (CF.R:194-325, 448-586): one evaluation on the device each ---------------------------------------------
## per-individual forms: `db` holds ONE individual of the resident training set (CF.R:655-657); mu_k_param must be resident
.need_resident <- function(mu_k_param, what)
  if (is.null(attr(mu_k_param, 'mcb_token')) || !identical(attr(mu_k_param, 'mcb_token'), .mcb$token))
    stop(what, ': mu_k_param is not the resident hyper-posterior; call e_step_VEM first or m_step_VEM / ',
         'logL_monitoring_VEM, which upload an explicit one')
.indiv_eval <- function(hp, db, mu_k_param, grad) {
  .need_resident(mu_k_param, 'logL_clust_multi_GP')
  r <- match(unique(db$ID)[1], .mcb$canon$ids)
  th <- .mcb$theta_i; th[, r] <- as.numeric(hp)[1:3]
  res <- .Call('mcbR_eval_indiv', mcb_session(), th, grad)
  if (grad) res[[2]][, r] else res[[1]][r]
}
logL_clust_multi_GP <- function(hp, db, mu_k_param, kern) { .check_kernels(kern); .indiv_eval(hp, db, mu_k_param, FALSE) }
gr_clust_multi_GP   <- function(hp, db, mu_k_param, kern) { .check_kernels(kern); .indiv_eval(hp, db, mu_k_param, TRUE) }

logL_clust_multi_GP_common_hp_i <- function(hp, db, mu_k_param, kern) {
  .check_kernels(kern); mcb_set_db(db); .need_resident(mu_k_param, 'logL_clust_multi_GP_common_hp_i')
  .Call('mcbR_eval_indiv_common', mcb_session(), as.numeric(hp)[1:3], FALSE)[1]
}
gr_clust_multi_GP_common_hp_i <- function(hp, db, mu_k_param, kern) {
  .check_kernels(kern); mcb_set_db(db); .need_resident(mu_k_param, 'gr_clust_multi_GP_common_hp_i')
  .Call('mcbR_eval_indiv_common', mcb_session(), as.numeric(hp)[1:3], TRUE)[2:4]
}

## mean-process forms on their explicit arguments (CF.R:194-216, 250-278, 448-475, 510-537): t, y, mean, new_cov are uploaded,
## nothing resident is trusted (n = length(t) is the grid size: one small upload per evaluation)
.gp_mod <- function(hp, t, y, mean, new_cov, pen_diag, grad) {
  res <- .Call('mcbR_gp_mod', mcb_session(), as.numeric(t), y, as.numeric(mean), new_cov, as.numeric(hp),
               ifelse(is.null(pen_diag), 0, pen_diag), grad)
  if (grad) res[-1] else res[1]
}
logL_GP_mod <- function(hp, db, mean, kern, new_cov, pen_diag = NULL) {
  .check_kernels(kern); .gp_mod(hp, db$Timestamp, as.numeric(db$Output), mean, as.numeric(new_cov), pen_diag, FALSE)
}
gr_GP_mod <- function(hp, db, mean, kern, new_cov, pen_diag = NULL) {
  .check_kernels(kern); .gp_mod(hp, db$Timestamp, as.numeric(db$Output), mean, as.numeric(new_cov), pen_diag, TRUE)
}
.gp_mod_common <- function(hp, db, mean, new_cov, pen_diag, grad) {
  names_k <- names(db); t <- db[[1]]$Timestamp                                             # t_k = db[[1]]$Timestamp, CF.R:264
  y <- sapply(names_k, function(k) as.numeric(db[[k]]$Output))                              # n x Z
  v <- lapply(names_k, function(k) as.numeric(mean[[k]]))
  m <- if (all(sapply(v, length) == 1)) unlist(v) else unlist(lapply(v, function(x) if (length(x) == 1) rep(x, length(t)) else x))
  cov <- unlist(lapply(names_k, function(k) as.numeric(new_cov[[k]])))
  .gp_mod(hp, t, y, m, cov, pen_diag, grad)
}
logL_GP_mod_common_hp_k <- function(hp, db, mean, kern, new_cov, pen_diag = NULL) {
  .check_kernels(kern); .gp_mod_common(hp, db, mean, new_cov, pen_diag, FALSE)
}
gr_GP_mod_common_hp_k <- function(hp, db, mean, kern, new_cov, pen_diag = NULL) {
  .check_kernels(kern); .gp_mod_common(hp, db, mean, new_cov, pen_diag, TRUE)
}

logL_monitoring_VEM <- function(hp, db, kern_i, kern_0, mu_k_param, m_k) {
  .check_kernels(kern_i, kern_0)
  cn <- mcb_set_db(db); names_k <- names(m_k)
  tok <- attr(mu_k_param, 'mcb_token')
  if (is.null(tok) || !identical(tok, .mcb$token)) .ensure_posterior(hp, m_k, mu_k_param, cn)
  .Call('mcbR_elbo', mcb_session(), .pack_theta_k(hp, names_k), .pack_theta_i(hp, cn$ids))[1]
}

BIC <- function(hp, db, kern_0, kern_i, m_k, mu_k_param, clust = TRUE) {
  if (!clust) stop('magmaclust_b200: BIC(clust = FALSE) needs the single-cluster model of the MAGMA repository')
  .check_kernels(kern_0, kern_i)
  cn <- mcb_set_db(db); names_k <- names(mu_k_param$mean); K <- length(names_k); N <- length(cn$all_t)
  attr(mu_k_param, 'mcb_token') <- NULL                                   # evaluate at exactly (hp, mu_k_param): upload both
  .ensure_posterior(hp, m_k, mu_k_param, cn)
  tr <- .Call('mcbR_bic_terms', mcb_session())                             # sum_i, sum_k, sum_k log det C_k   (CF.R:363-420)
  pen <- 0.5 * (dplyr::n_distinct(unlist(hp$theta_i)) + dplyr::n_distinct(unlist(hp$theta_k)) +
                dplyr::n_distinct(unlist(hp$pi_k)) - 1) * log(length(cn$ids))
  tr[1] + tr[2] + 0.5 * (K * (N + N * log(2 * pi)) + tr[3]) - pen
}

## ---- drivers (MC.R) ------------------------------------------------------------------------------------------------------------
training_VEM <- function(db, prior_mean_k, ini_hp = list('theta_k' = c(1, 1, 0.2), 'theta_i' = c(1, 1, 0.2)),
                         kern_0 = kernel_mu, kern_i = kernel, ini_tau_i_k = NULL, common_hp_k = T, common_hp_i = T) {
  .check_kernels(kern_0, kern_i)
  n_loop_max <- 15
  cn <- mcb_set_db(db); names_k <- names(prior_mean_k); K <- length(names_k); M <- length(cn$ids); N <- length(cn$all_t)
  if (is.null(ini_tau_i_k)) ini_tau_i_k <- ini_tau_i_k(db, k = K, nstart = 50)              # MC.R:37 (host k-means)
  tau0 <- .pack_tau(ini_tau_i_k, names_k, cn$ids)
  ## MC.R:39: pi_k in the LIST order of ini_tau_i_k, multiplied positionally by update_tau_i_k_VEM (CF.R:757)
  pi_k <- sapply(ini_tau_i_k, function(x) mean(unlist(x)))
  t1 <- Sys.time()
  .set_state(matrix(as.numeric(ini_hp$theta_k)[1:3], 3, K), matrix(as.numeric(ini_hp$theta_i)[1:3], 3, M), pi_k,
             .pack_mk(prior_mean_k, names_k, N), tau0)
  seq_logL <- c(); cv <- 'FALSE'; it <- NULL
  for (i in 1:n_loop_max) {
    ## E-step, ELBO, M-step, stopping rule of MC.R:44-93 in one call; `param` of the E-step is copied out behind the M-step
    it <- .Call('mcbR_vem_iteration_host', mcb_session(), common_hp_k, common_hp_i, K, M, N)
    if (!is.finite(it[[1]][2])) {                                                           # M-step failure, MC.R:62-67
      print(paste0('The M-step encountered an error at iteration : ', i)); break
    }
    seq_logL <- c(seq_logL, it[[1]][1])
    if (it[[1]][3] != 0) { cv <- TRUE; break }
  }
  t2 <- Sys.time()
  nh <- .Call('mcbR_get_new_hp', mcb_session(), K, M)                                       # hp = new_hp, MC.R:95
  param <- .param_out(it[[2]], it[[3]], it[[4]], names_k, cn)
  hp <- list(theta_k = setNames(lapply(seq_len(K), function(k) nh[[1]][, k]), names_k),
             theta_i = setNames(lapply(seq_len(M), function(i) nh[[2]][, i]), cn$ids), pi_k = setNames(nh[[3]], names_k))
  list('hp' = hp, 'convergence' = cv, 'param' = param, 'Time_train' = difftime(t2, t1, units = 'secs'),
       'logL_iterations' = seq_logL)
}

posterior_mu_k <- function(db, timestamps, m_k, kern_0, kern_i, list_hp) {
  .check_kernels(kern_0, kern_i)
  cn <- mcb_set_db(db); names_k <- names(list_hp$hp$theta_k); K <- length(names_k)
  tau <- .pack_tau(list_hp$param$tau_i_k, names_k, cn$ids)
  pi_k <- if (is.null(list_hp$hp$pi_k)) rowMeans(tau) else list_hp$hp$pi_k
  .set_state(.pack_theta_k(list_hp$hp, names_k), .pack_theta_i(list_hp$hp, cn$ids), pi_k,
             .pack_mk(m_k, names_k, length(cn$all_t)), tau)
  timestamps <- sort(unique(timestamps))
  res <- .Call('mcbR_posterior_mu_k', mcb_session(), as.numeric(timestamps), tau, K)
  nm <- paste0('X', timestamps)
  mean_k <- lapply(seq_len(K), function(k) tibble::tibble(Timestamp = timestamps, Output = res[[1]][, k]))
  cov_k <- lapply(seq_len(K), function(k) { m <- res[[2]][, , k]; dimnames(m) <- list(nm, nm); m })
  out <- list('mean' = setNames(mean_k, names_k), 'cov' = setNames(cov_k, names_k), 'tau_i_k' = list_hp$param$tau_i_k)
  .mcb$pred_token <- .new_token(); attr(out, 'mcb_pred_token') <- .mcb$pred_token
  out
}

## make list(mean, cov) the resident prediction posterior (uploaded unless it is the last posterior_mu_k's result)
.ensure_pred <- function(mean_k, cov_k, tok) {
  grid <- mean_k[[1]]$Timestamp
  if (!is.null(tok) && identical(tok, .mcb$pred_token)) return(grid)
  names_k <- names(mean_k)
  mean <- sapply(names_k, function(k) as.numeric(mean_k[[k]]$Output))                        # T x K
  cov <- array(unlist(lapply(names_k, function(k) as.numeric(cov_k[[k]]))), c(length(grid), length(grid), length(names_k)))
  .Call('mcbR_set_pred_posterior', mcb_session(), as.numeric(grid), mean, cov, NULL)
  .mcb$pred_token <- NULL
  grid
}
.grid_pos <- function(t, grid, what) {
  p <- match(t, grid)
  if (anyNA(p)) stop(what, ': timestamps missing from the grid of the mean processes')
  as.integer(p - 1L)
}

update_tau_star_k_EM <- function(db, mean_k, cov_k, kern, hp, pi_k) {
  .check_kernels(kern)
  grid <- .ensure_pred(mean_k, cov_k, attr(mean_k, 'mcb_pred_token'))
  names_k <- names(mean_k)
  res <- .Call('mcbR_predict', mcb_session(), c(0L, nrow(db)), .grid_pos(db$Timestamp, grid, 'update_tau_star_k_EM'),
               as.numeric(db$Output), as.numeric(hp)[1:3], as.numeric(unlist(pi_k)), integer(0), length(names_k))
  split(as.vector(res[[1]][, 1]), names_k)
}

train_new_gp_EM <- function(db, param_mu_k, ini_hp = NULL, kern_i, hp_i = NULL) {
  .check_kernels(kern_i)
  mean_mu_k <- param_mu_k$mean; cov_mu_k <- param_mu_k$cov
  attr(mean_mu_k, 'mcb_pred_token') <- attr(param_mu_k, 'mcb_pred_token')
  pi_k <- lapply(param_mu_k$tau_i_k, function(x) Reduce('+', x) / length(x))                # MC.R:141
  if (is.null(hp_i)) {
    ## free-hp branch, MC.R:142-186 (repaired, DESIGN.md): the whole EM loop runs in the library
    grid <- .ensure_pred(mean_mu_k, cov_mu_k, attr(mean_mu_k, 'mcb_pred_token'))
    res <- .Call('mcbR_train_new_indiv', mcb_session(), c(0L, nrow(db)), .grid_pos(db$Timestamp, grid, 'train_new_gp_EM'),
                 as.numeric(db$Output), as.numeric(ini_hp$theta_i)[1:3], as.numeric(unlist(pi_k)), 25L)
    new_hp <- res[[1]][, 1]
    tau_k <- split(as.vector(res[[2]][, 1]), names(mean_mu_k))
  } else {
    new_hp <- hp_i$hp$theta_i[[1]]                                                          # MC.R:189
    tau_k <- update_tau_star_k_EM(db, mean_mu_k, cov_mu_k, kern_i, new_hp, pi_k)
  }
  list('theta_new' = new_hp, 'tau_k' = tau_k)
}

pred_gp_clust <- function(db, timestamps = NULL, list_mu, kern, hp) {
  .check_kernels(kern)
  tn <- db$Timestamp
  if (is.null(timestamps)) timestamps <- seq(min(tn), max(tn), length.out = 500)
  mean_k <- list_mu$mean
  grid <- .ensure_pred(mean_k, list_mu$cov, attr(list_mu, 'mcb_pred_token'))
  names_k <- names(mean_k)
  ## tau_k is an input here (hp$tau_k, MC.R:252,269); means and variances for all clusters in one launch
  res <- .Call('mcbR_predict', mcb_session(), c(0L, nrow(db)), .grid_pos(tn, grid, 'pred_gp_clust'), as.numeric(db$Output),
               as.numeric(hp$theta_new)[1:3], NULL, .grid_pos(timestamps, grid, 'pred_gp_clust'), length(names_k))
  floop <- function(k) tibble::tibble('Timestamp' = timestamps, 'Mean' = res[[2]][, k, 1], 'Var' = res[[3]][, k, 1],
                                      'tau_k' = hp$tau_k[[names_k[k]]])
  setNames(lapply(seq_along(names_k), floop), names_k)
}

full_algo_clust <- function(db, new_db, timestamps, kern_i, ini_tau_i_k = NULL, common_hp_k = T, common_hp_i = T,
                            prior_mean = NULL, kern_0 = NULL, list_hp = NULL, mu_k = NULL, ini_hp = NULL, hp_new_i = NULL) {
  if (is.null(list_hp))
    list_hp <- training_VEM(db, prior_mean, ini_hp, kern_0, kern_i, ini_tau_i_k, common_hp_k, common_hp_i)
  t_pred <- sort(union(union(timestamps, unique(db$Timestamp)), unique(new_db$Timestamp)))
  if (is.null(mu_k)) mu_k <- posterior_mu_k(db, t_pred, prior_mean, kern_0, kern_i, list_hp)
  if (is.null(hp_new_i)) {
    hp_new_i <- if (common_hp_i) train_new_gp_EM(new_db, mu_k, ini_hp, kern_i, hp_i = list_hp)
                else train_new_gp_EM(new_db, mu_k, ini_hp, kern_i)
  }
  pred <- pred_gp_clust(new_db, timestamps, mu_k, kern_i, hp_new_i)
  list('Prediction' = pred, 'Mean_processes' = mu_k, 'Hyperparameters' = list('other_i' = list_hp, 'new_i' = hp_new_i),
       'logL_iterations' = list_hp$logL_iterations, 'ari_iterations' = list_hp$ARI_iterations)
}

pred_max_cluster <- function(tau_i_k) apply(sapply(tau_i_k, unlist), 1, which.max)          # MC.R:349-355: first maximum

model_selection <- function(db, k_grid = 2:5, ini_hp = list('theta_k' = c(1, 1), 'theta_i' = c(1, 1, 0.2)),
                            kern_0 = kernel_mu, kern_i = kernel, ini_tau_i_k = NULL, common_hp_k = T, common_hp_i = T,
                            plot = F) {
  floop <- function(K) {
    if (K == 1) stop('magmaclust_b200: K = 1 needs training() of the MAGMA repository (MC.R:113 calls an undefined function)')
    prior_mean_k <- as.list(setNames(rep(0, K), paste0('K', seq_len(K))))
    ini <- if (is.null(ini_tau_i_k)) get('ini_tau_i_k', envir = globalenv())(db, K) else ini_tau_i_k
    hp0 <- ini_hp; if (length(hp0$theta_k) == 2) hp0$theta_k <- c(hp0$theta_k, 0.2)
    model <- training_VEM(db, prior_mean_k, hp0, kern_0, kern_i, ini, common_hp_k, common_hp_i)
    model[['BIC']] <- BIC(model$hp, db, kern_0, kern_i, prior_mean_k, model$param, TRUE)
    model
  }
  res <- setNames(lapply(k_grid, floop), paste0('K = ', k_grid))
  bic <- sapply(res, function(m) m$BIC)
  res$K_max_BIC <- k_grid[which.max(bic)]
  res
}

mcb_version <- function() .Call('mcbR_version')
