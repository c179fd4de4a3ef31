## magmaclust_b200.R — drop-in definitions of the reference's hot-path functions on top of the .Call shim.
##
## SOURCE ONLY (never run: R is not installed in the build image). Usage in the reference:
##   source('Computing_functions.R'); source('r/R/magmaclust_b200.R')   # the definitions below mask the R ones
## Signatures and return shapes are those of Computing_functions.R / MAGMAclust.R; see INTEGRATION.md.

.mcb <- new.env()

mcb_session <- function(device = 0L) {
  if (is.null(.mcb$h)) {
    dyn.load('r/src/mcb_shim.so')
    .mcb$h <- .Call('mcbR_create', as.integer(device))
  }
  .mcb$h
}

## CSR canonicalisation: replaces db %>% filter(ID == i) and the 'X<t>' dimnames (CF.R:83,146,601)
mcb_set_db <- function(db) {
  key <- digest_db(db)
  if (identical(.mcb$db_key, key)) return(invisible(.mcb$canon))
  ids <- unique(db$ID)
  all_t <- sort(unique(db$Timestamp))
  ord <- order(match(db$ID, ids))                         # blocks in unique(db$ID) order, rows kept
  counts <- as.integer(table(factor(db$ID, levels = ids)))
  canon <- list(ids = ids, all_t = all_t, offsets = c(0L, cumsum(counts)),
                grid_idx = match(db$Timestamp[ord], all_t) - 1L, y = db$Output[ord])
  .Call('mcbR_set_data', mcb_session(), canon$offsets, canon$grid_idx, canon$y, all_t)
  .mcb$db_key <- key; .mcb$canon <- canon
  invisible(canon)
}
digest_db <- function(db) c(nrow(db), sum(db$Timestamp), sum(db$Output))

.check_kernels <- function(...) for (k in list(...))
  if (!identical(k, kernel) && !identical(k, kernel_mu))
    stop('magmaclust_b200: only the squared-exponential kernels kernel / kernel_mu are implemented on the GPU')

e_step_VEM <- function(db, m_k, kern_0, kern_i, hp, old_tau_i_k) {
  .check_kernels(kern_0, kern_i)
  cn <- mcb_set_db(db); K <- length(m_k); names_k <- names(m_k)
  theta_k <- sapply(names_k, function(k) as.numeric(hp$theta_k[[k]])[1:3])            # 3 x K
  theta_i <- sapply(as.character(cn$ids), function(i) as.numeric(unlist(hp$theta_i[[i]]))[1:3])
  tau <- sapply(as.character(cn$ids), function(i) sapply(names_k, function(k) old_tau_i_k[[k]][[i]]))  # K x M
  res <- .Call('mcbR_e_step_VEM', mcb_session(), theta_k, theta_i, as.numeric(hp$pi_k), as.numeric(unlist(m_k)),
               tau, length(cn$all_t))
  nm <- paste0('X', cn$all_t)
  mean_k <- lapply(seq_len(K), function(k) tibble::tibble(Timestamp = cn$all_t, Output = res[[1]][, k]))
  cov_k <- lapply(seq_len(K), function(k) { m <- res[[2]][, , k]; dimnames(m) <- list(nm, nm); m })
  tau_k <- lapply(seq_len(K), function(k) setNames(as.list(res[[3]][k, ]), cn$ids))
  list(mean = setNames(mean_k, names_k), cov = setNames(cov_k, names_k), tau_i_k = setNames(tau_k, names_k))
}

## optimiser callbacks keep their signatures; the heavy arguments (db, mu_k_param, new_cov) are the resident ones
logL_clust_multi_GP_common_hp_i <- function(hp, db, mu_k_param, kern) .Call('mcbR_eval_indiv_common', mcb_session(), as.numeric(hp), FALSE)[1]
gr_clust_multi_GP_common_hp_i   <- function(hp, db, mu_k_param, kern) .Call('mcbR_eval_indiv_common', mcb_session(), as.numeric(hp), TRUE)[2:4]
logL_GP_mod_common_hp_k <- function(hp, db, mean, kern, new_cov, pen_diag = NULL)
  .Call('mcbR_eval_mean', mcb_session(), -1L, as.numeric(hp), ifelse(is.null(pen_diag), 0, pen_diag), FALSE)[1]
gr_GP_mod_common_hp_k <- function(hp, db, mean, kern, new_cov, pen_diag = NULL)
  .Call('mcbR_eval_mean', mcb_session(), -1L, as.numeric(hp), ifelse(is.null(pen_diag), 0, pen_diag), TRUE)[-1]

## m_step_VEM with the library's own batched L-BFGS (one kernel for all per-individual optimisations)
m_step_VEM <- function(db, old_hp, list_mu_param, kern_0, kern_i, m_k, common_hp_k, common_hp_i) {
  .check_kernels(kern_0, kern_i)
  cn <- mcb_set_db(db); names_k <- names(m_k)
  res <- .Call('mcbR_m_step', mcb_session(), common_hp_k, common_hp_i, length(m_k), length(cn$ids))
  theta_k <- lapply(seq_along(names_k), function(k) res[[1]][, k])
  theta_i <- lapply(seq_along(cn$ids), function(i) res[[2]][, i])
  list(theta_k = setNames(theta_k, names_k), theta_i = setNames(theta_i, cn$ids), pi_k = setNames(res[[3]], names_k))
}

## the building blocks the reference also calls directly (vector branches only; the db branches loop over the IDs in R)
kern_to_cov <- function(x, kern = kernel, theta = c(1, 0.5), sigma = 0.2) {
  .check_kernels(kern)
  if (is.vector(x)) {
    x <- sort(x)
    m <- .Call('mcbR_kern', mcb_session(), as.numeric(x), as.numeric(theta)[1:2], sigma, FALSE)
    dimnames(m) <- list(paste0('X', x), paste0('X', x)); m
  } else {
    floop <- function(i) { th <- if (is.list(theta)) theta[[i]] else c(theta, sigma)
      kern_to_cov(x$Timestamp[x$ID == i], kern, th[1:2], th[[3]]) }
    res <- sapply(unique(x$ID), floop, simplify = FALSE, USE.NAMES = TRUE)
    if (length(res) == 1) res[[1]] else res
  }
}
kern_to_inv <- function(x, kern = kernel, theta = c(1, 0.5), sigma = 0.2) {
  .check_kernels(kern)
  if (is.vector(x)) {
    x <- sort(x)
    m <- .Call('mcbR_kern', mcb_session(), as.numeric(x), as.numeric(theta)[1:2], sigma, TRUE)
    attr(m, 'logdet') <- NULL
    dimnames(m) <- list(paste0('X', x), paste0('X', x)); m
  } else {
    floop <- function(i) { th <- if (is.list(theta)) theta[[i]] else c(theta, sigma)
      kern_to_inv(x$Timestamp[x$ID == i], kern, th[1:2], th[[3]]) }
    res <- sapply(unique(x$ID), floop, simplify = FALSE, USE.NAMES = TRUE)
    if (length(res) == 1) res[[1]] else res
  }
}
dmvnorm <- function(x, mu, inv_Sigma, log = FALSE)
  .Call('mcbR_dmvnorm', mcb_session(), as.numeric(x), as.numeric(mu), as.numeric(inv_Sigma), log)

