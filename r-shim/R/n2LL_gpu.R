# Drop-in replacement of the objective inside varycoef (R-pkg/R/objective_functions.R:4-62).
# Same formals as n2LL, so optim(fn = n2LL, ...) in MLE_computation (R-pkg/R/SVC_mle.R:145-163),
# optimParallel (:180-199) and SVC_selection's do.call(obj.fun$obj_fun, c(list(x = x), obj.fun$args))
# (R-pkg/R/SVC_selection.R:74-76) keep working unchanged.  The heavy closure arguments are replaced
# by `gpu`, an external pointer created once per fit.

vcb_handle <- function(y, X, locs, W, control) {
  cov_id   <- match(control$cov.name, c("exp", "mat32", "mat52", "sph", "wend1", "wend2")) - 1L
  dist_id  <- match(control$dist$method, c("euclidean", "maximum", "manhattan", "minkowski")) - 1L
  taper_id <- if (is.null(control$tapering)) 0L else
    switch(control$cov.name, exp = 1L, mat32 = 1L, mat52 = 2L,
           stop("tapering undefined for this covariance"))       # get_taper, utils.R:545-552
  locs <- as.matrix(locs); X <- as.matrix(X); W <- as.matrix(W)
  storage.mode(locs) <- storage.mode(X) <- storage.mode(W) <- "double"   # integer grids / dummies
  .Call(C_vcb_create, locs, X, W, as.numeric(y),
        cov_id, dist_id, if (is.null(control$dist$p)) 2 else control$dist$p,
        taper_id, if (is.null(control$tapering)) 0 else control$tapering, 0L)
}

# control$parallel = integer vector of GPU ordinals -> one replica per GPU; batches (optim's stencils, submitted through
# .Call(C_vcb_n2ll_batch, gpu, thetas, mus)) are spread over them: replaces the optimParallel cluster of SVC_mle.R:164-200
vcb_handle_multi <- function(y, X, locs, W, control, devices) {
  cov_id   <- match(control$cov.name, c("exp", "mat32", "mat52", "sph", "wend1", "wend2")) - 1L
  dist_id  <- match(control$dist$method, c("euclidean", "maximum", "manhattan", "minkowski")) - 1L
  taper_id <- if (is.null(control$tapering)) 0L else switch(control$cov.name, exp = 1L, mat32 = 1L, mat52 = 2L)
  .Call(C_vcb_create_multi, as.matrix(locs), as.matrix(X), as.matrix(W), as.numeric(y), cov_id, dist_id,
        if (is.null(control$dist$p)) 2 else control$dist$p, taper_id,
        if (is.null(control$tapering)) 0 else control$tapering, as.integer(devices))
}

# med_dist <- median(as.numeric(d)) (SVC_mle.R:39-45) without the n x n matrix
med_dist_gpu <- function(locs, control)
  .Call(C_vcb_median_dist, as.matrix(locs),
        match(control$dist$method, c("euclidean", "maximum", "manhattan", "minkowski")) - 1L,
        if (is.null(control$dist$p)) 2 else control$dist$p, 0L)

n2LL <- function(x, cov_func, outer.W, y, X, W, mean.est = NULL, taper = NULL,
                 pc.dens = NULL, Rstruct = NULL, profile = TRUE, gpu = NULL) {
  q <- dim(W)[2]; p <- dim(X)[2]
  mu <- if (profile) mean.est else x[1 + 2*q + 1:p]     # NULL -> GLS on the GPU
  as.numeric(.Call(C_vcb_n2ll, gpu, x[1:(2*q+1)], mu, p)) + pc_penalty(x, q, pc.dens)
}

# In MLE_computation (R-pkg/R/SVC_mle.R:12-242) the edits are:
#   - drop `d <- own_dist(...)`, `outer.W <- lapply(...)`, `taper`, `Rstruct`   (lines 33-36, 63-81)
#   - gpu <- vcb_handle(y, X, locs, W, control)
#   - pass `gpu = gpu` in the optim()/optimParallel() argument list and in `args` of extract_fun
#   - med_dist <- med_dist_gpu(locs, control)   (dense branch; the tapered branch is the constant 1, SVC_mle.R:44)

# After the optimisation (R-pkg/R/SVC_mle.R:205-296) the O(n^3) post-processing also comes from the resident factor:
#   post <- .Call(C_vcb_post, gpu, cov.par, p, TRUE)      # list(mu_GLS, Sigma_FE = (X' S^-1 X)^-1, edof)
#     replaces solve(Sigma_final) in prep_par_output (utils.R:423-473) and eff_dof() (eff_dof.R:7-21)
#   pr <- .Call(C_vcb_predict, gpu, cov.par, mu, newlocs, newX, newW, q, compute.y.var)
#     replaces the body of predict.SVC_mle (predict-SVC_mle.R:80-267): list(eff, y.pred, y.var)
# SVC_selection: CD_mu's  R.t.inv <- solve(t(R)); X.tilde <- R.t.inv %*% X  (SVC_selection.R:30-35) becomes
#   .Call(C_vcb_n2ll, gpu, theta.k, rep(0, p), p); Z <- .Call(C_vcb_whitened, gpu, n, p)   # cbind(X.tilde, y.tilde)
GLS_chol.matrix <- function(R, X, y) {                  # p x 1 matrix, as utils.R:93-101 returns it
  X <- as.matrix(X); storage.mode(X) <- "double"; storage.mode(R) <- "double"
  .Call(C_vcb_gls_chol, R, X, as.numeric(y), 0L)
}
