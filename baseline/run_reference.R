## run_reference.R — times the UNTOUCHED reference (ArthurLeroy/MAGMAclust) on the files data/gen.py wrote.
##
##   Rscript baseline/run_reference.R <reference dir> <data dir, e.g. data/generated/C1> [max individuals] [gpu]
##
## NOT RUN in this repository's build image or on the GPU box: neither has R (no R, Rscript, libR.so; SURVEY.md §8c/§8d).
## It is shipped so that a maintainer with R + tidyverse + MASS + optimr can produce the reference column of the comparison
## on the same bytes bench.py reads (bench.py --data <data dir>). With the 4th argument `gpu` the same script runs the same
## calls with r/R/magmaclust_b200.R sourced on top (the drop-in), and prints the same table.
##
## What is timed, with Sys.time() brackets exactly like the reference's own (MAGMAclust.R:41, 94-96):
##   e_step_VEM, logL_monitoring_VEM, m_step_VEM                   one loop body of training_VEM from the study's start state
##   training_VEM                                                  whole training (<= 15 iterations)
##   posterior_mu_k, train_new_gp_EM (fixed hp), pred_gp_clust     prediction of the first individual as if it were new
## The reference is single-threaded R (Simulation_study.R:543); BLAS threads are whatever R links (printed below).

args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 2) stop('usage: Rscript run_reference.R <reference dir> <data dir> [max individuals] [gpu]')
ref_dir <- normalizePath(args[1]); data_dir <- normalizePath(args[2])
max_m <- if (length(args) >= 3 && args[3] != 'all') as.integer(args[3]) else NA
use_gpu <- length(args) >= 4 && args[4] == 'gpu'
repo <- normalizePath(file.path(dirname(sub('--file=', '', grep('--file=', commandArgs(), value = TRUE)[1])), '..'))

old <- setwd(ref_dir)
suppressMessages(source('MAGMAclust.R'))            # sources Computing_functions.R itself (MAGMAclust.R:13)
setwd(old)
if (use_gpu) { setwd(repo); source('r/R/magmaclust_b200.R') }

meta <- jsonlite::fromJSON(file.path(data_dir, 'meta.json'))
db <- readr::read_csv(file.path(data_dir, 'db.csv'), col_types = 'cddc') %>% dplyr::select(ID, Timestamp, Output)
tau0 <- readr::read_csv(file.path(data_dir, 'tau0.csv'), col_types = readr::cols(ID = 'c', .default = 'd'))
if (!is.na(max_m)) {                                 # a bounded sample: the first max_m individuals (state it in the report)
  keep <- unique(db$ID)[seq_len(min(max_m, dplyr::n_distinct(db$ID)))]
  db <- db %>% dplyr::filter(ID %in% keep); tau0 <- tau0 %>% dplyr::filter(ID %in% keep)
}
K <- meta$K
names_k <- paste0('K', seq_len(K))
prior_mean_k <- as.list(setNames(rep(meta$start_state$m_k, K), names_k))
ini_tau <- setNames(lapply(names_k, function(k) setNames(as.list(tau0[[k]]), tau0$ID)), names_k)
common_hp_i <- isTRUE(meta$config$common_hp_i)
ids <- unique(db$ID)
hp <- list(theta_k = setNames(rep(list(meta$start_state$theta_k), K), names_k),
           theta_i = setNames(rep(list(meta$start_state$theta_i), length(ids)), ids),
           pi_k = sapply(ini_tau, function(x) mean(unlist(x))))

timed <- function(label, expr) {
  t1 <- Sys.time(); v <- force(expr); t2 <- Sys.time()
  cat(sprintf('%-28s %10.3f s\n', label, as.numeric(difftime(t2, t1, units = 'secs'))))
  invisible(v)
}
cat(sprintf('config %s  M = %d  K = %d  N = %d  common_hp_i = %s  impl = %s  R %s  BLAS %s\n', meta$config$name, length(ids), K,
            dplyr::n_distinct(db$Timestamp), common_hp_i, ifelse(use_gpu, 'magmaclust_b200', 'reference'),
            getRversion(), tryCatch(extSoftVersion()[['BLAS']], error = function(e) 'unknown')))

## one loop body of training_VEM (MC.R:44-93)
param <- timed('e_step_VEM', e_step_VEM(db, prior_mean_k, kernel_mu, kernel, hp, ini_tau))
elbo <- timed('logL_monitoring_VEM', logL_monitoring_VEM(hp, db, kernel, kernel_mu, mu_k_param = param, m_k = prior_mean_k))
new_hp <- timed('m_step_VEM', m_step_VEM(db, hp, list_mu_param = param, kernel_mu, kernel, prior_mean_k, TRUE, common_hp_i))
cat(sprintf('ELBO at the start state %.10g\n', elbo))

## whole training and the prediction chain (MC.R:301-347 without the plotting)
model <- timed('training_VEM', training_VEM(db, prior_mean_k, list(theta_k = meta$start_state$theta_k, theta_i = meta$start_state$theta_i),
                                            kernel_mu, kernel, ini_tau, TRUE, common_hp_i))
cat(sprintf('iterations %d  converged %s  Time_train %.3f s  last ELBO %.10g\n', length(model$logL_iterations),
            model$convergence, as.numeric(model$Time_train), tail(model$logL_iterations, 1)))
new_db <- db %>% dplyr::filter(ID == ids[1]) %>% dplyr::select(Timestamp, Output)
t_pred <- sort(unique(db$Timestamp))
mu_k <- timed('posterior_mu_k', posterior_mu_k(db, t_pred, prior_mean_k, kernel_mu, kernel, model))
hp_new <- timed('train_new_gp_EM (fixed hp)', train_new_gp_EM(new_db, mu_k, NULL, kernel, hp_i = model))
pred <- timed('pred_gp_clust', pred_gp_clust(new_db, t_pred, mu_k, kernel, hp_new))
cat('hard assignments: ', paste(pred_max_cluster(model$param$tau_i_k), collapse = ' '), '\n')
