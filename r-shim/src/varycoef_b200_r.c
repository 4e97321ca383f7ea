/*
 * .Call shim between varycoef's R host code and libvarycoef_b200.so (include/varycoef_b200.h).
 * Not compiled in this repository's image (no R headers here); build inside varycoef with
 *   PKG_CPPFLAGS = -I<repo>/include     PKG_LIBS = -L<repo>/varycoef_b200 -lvarycoef_b200
 * and register with useDynLib(varycoef, .registration = TRUE).
 *
 * Replaces, on the R side: own_dist + cov.func + outer.W set-up of MLE_computation
 * (R-pkg/R/SVC_mle.R:33-72) and the body of n2LL (R-pkg/R/objective_functions.R:4-62).
 */
#include <R.h>
#include <Rinternals.h>
#include "varycoef_b200.h"

static void vcb_finalizer(SEXP ext) {
  vc_handle* h = (vc_handle*) R_ExternalPtrAddr(ext);
  if (h) { vc_destroy(h); R_ClearExternalPtr(ext); }
}

static vc_handle* vcb_get(SEXP ext) {
  vc_handle* h = (vc_handle*) R_ExternalPtrAddr(ext);
  if (!h) Rf_error("varycoef_b200: handle was destroyed or restored from a serialised session");
  return h;
}

static void vcb_check(int rc, vc_handle* h) {
  if (rc == VC_OK) return;
  if (rc == VC_ERR_NOT_PD)   /* same wording as base::chol */
    Rf_error("the leading minor of order %d is not positive", vc_last_info(h));
  Rf_error("varycoef_b200 (%d): %s", rc, vc_last_error(h));
}

/* .Call(C_vcb_create, locs, X, W, y, cov_id, dist_id, mink_p, taper_id, taper_range, device) */
SEXP vcb_create(SEXP locs, SEXP X, SEXP W, SEXP y, SEXP cov_id, SEXP dist_id, SEXP mink_p,
                SEXP taper_id, SEXP taper_range, SEXP device) {
  vc_config cfg;
  vc_config_default(&cfg);
  cfg.cov_id = Rf_asInteger(cov_id);
  cfg.dist_method = Rf_asInteger(dist_id);
  cfg.minkowski_p = Rf_asReal(mink_p);
  cfg.taper_id = Rf_asInteger(taper_id);
  cfg.taper_range = Rf_asReal(taper_range);
  cfg.device = Rf_asInteger(device);
  /* integer coordinates / design columns (a 1:n grid, dummy variables) arrive as INTSXP: coerce, REAL() would fail */
  locs = PROTECT(Rf_coerceVector(locs, REALSXP));
  X = PROTECT(Rf_coerceVector(X, REALSXP));
  W = PROTECT(Rf_coerceVector(W, REALSXP));
  y = PROTECT(Rf_coerceVector(y, REALSXP));
  const int64_t n = Rf_nrows(X);
  vc_handle* h = NULL;
  int rc = vc_create(&h, &cfg, n, Rf_ncols(locs), Rf_ncols(X), Rf_ncols(W),
                     REAL(locs), REAL(X), REAL(W), REAL(y));
  if (rc != VC_OK) Rf_error("varycoef_b200 (%d): %s", rc, vc_last_error(NULL));
  SEXP ext = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ext, vcb_finalizer, TRUE);
  UNPROTECT(5);
  return ext;
}

/* .Call(C_vcb_create_multi, locs, X, W, y, cov_id, dist_id, mink_p, taper_id, taper_range, devices): one replica per GPU;
 * vcb_n2ll_batch then spreads a batch over them -- what control$parallel / optimParallel did with PSOCK workers
 * (R-pkg/R/SVC_mle.R:164-200) */
SEXP vcb_create_multi(SEXP locs, SEXP X, SEXP W, SEXP y, SEXP cov_id, SEXP dist_id, SEXP mink_p,
                      SEXP taper_id, SEXP taper_range, SEXP devices) {
  vc_config cfg;
  vc_config_default(&cfg);
  cfg.cov_id = Rf_asInteger(cov_id);
  cfg.dist_method = Rf_asInteger(dist_id);
  cfg.minkowski_p = Rf_asReal(mink_p);
  cfg.taper_id = Rf_asInteger(taper_id);
  cfg.taper_range = Rf_asReal(taper_range);
  locs = PROTECT(Rf_coerceVector(locs, REALSXP));
  X = PROTECT(Rf_coerceVector(X, REALSXP));
  W = PROTECT(Rf_coerceVector(W, REALSXP));
  y = PROTECT(Rf_coerceVector(y, REALSXP));
  devices = PROTECT(Rf_coerceVector(devices, INTSXP));
  vc_handle* h = NULL;
  int rc = vc_create_multi(&h, &cfg, (int64_t) Rf_nrows(X), Rf_ncols(locs), Rf_ncols(X), Rf_ncols(W),
                           REAL(locs), REAL(X), REAL(W), REAL(y), INTEGER(devices), Rf_length(devices));
  if (rc != VC_OK) Rf_error("varycoef_b200 (%d): %s", rc, vc_last_error(NULL));
  SEXP ext = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ext, vcb_finalizer, TRUE);
  UNPROTECT(6);
  return ext;
}

/* .Call(C_vcb_median_dist, locs, dist_id, mink_p, device) -> med_dist: median(as.numeric(d)) of MLE_computation
 * (R-pkg/R/SVC_mle.R:39-45, dense branch) without the n x n distance matrix */
SEXP vcb_median_dist(SEXP locs, SEXP dist_id, SEXP mink_p, SEXP device) {
  locs = PROTECT(Rf_coerceVector(locs, REALSXP));
  double med = NA_REAL;
  int rc = vc_median_dist(Rf_asInteger(device), (int64_t) Rf_nrows(locs), Rf_ncols(locs), REAL(locs),
                          Rf_asInteger(dist_id), Rf_asReal(mink_p), &med, NULL);
  UNPROTECT(1);
  if (rc != VC_OK) Rf_error("varycoef_b200 (%d): %s", rc, vc_last_error(NULL));
  return Rf_ScalarReal(med);
}

/* .Call(C_vcb_n2ll, handle, theta, mu or NULL, p) -> numeric(1) with attributes mu (length p), logdet, quad */
SEXP vcb_n2ll(SEXP ext, SEXP theta, SEXP mu, SEXP p_) {
  vc_handle* h = vcb_get(ext);
  const int fixed = !Rf_isNull(mu);
  const int p = Rf_asInteger(p_);
  if (p < 1 || (fixed && Rf_length(mu) != p)) Rf_error("varycoef_b200: mu must have length p");
  double val, logdet, quad;
  SEXP mu_out = PROTECT(Rf_allocVector(REALSXP, p));
  int rc = vc_n2ll(h, REAL(theta), fixed ? VC_MU_FIXED : VC_MU_GLS, fixed ? REAL(mu) : NULL,
                   &val, REAL(mu_out), &logdet, &quad);
  vcb_check(rc, h);
  SEXP out = PROTECT(Rf_ScalarReal(val));
  Rf_setAttrib(out, Rf_install("mu"), mu_out);       /* the mean used: GLS estimate or the caller's mu */
  Rf_setAttrib(out, Rf_install("logdet"), Rf_ScalarReal(logdet));
  Rf_setAttrib(out, Rf_install("quad"), Rf_ScalarReal(quad));
  UNPROTECT(2);
  return out;
}

/* .Call(C_vcb_n2ll_batch, handle, thetas[(2q+1) x m], mus[p x m] or NULL) -> numeric(m) */
SEXP vcb_n2ll_batch(SEXP ext, SEXP thetas, SEXP mus) {
  vc_handle* h = vcb_get(ext);
  const int m = Rf_ncols(thetas);
  const int fixed = !Rf_isNull(mus);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, m));
  int rc = vc_n2ll_batch(h, m, REAL(thetas), fixed ? VC_MU_FIXED : VC_MU_GLS,
                         fixed ? REAL(mus) : NULL, REAL(out), NULL, NULL);
  vcb_check(rc, h);
  UNPROTECT(1);
  return out;
}

/* .Call(C_vcb_post, handle, theta, p, want_edof) -> list(mu, Sigma_FE, edof): the O(n^3) pieces of
 * prep_par_output (R-pkg/R/utils.R:423-473) and eff_dof (R-pkg/R/eff_dof.R:7-21) from the resident factor */
SEXP vcb_post(SEXP ext, SEXP theta, SEXP p_, SEXP want_edof) {
  vc_handle* h = vcb_get(ext);
  const int p = Rf_asInteger(p_);
  SEXP mu = PROTECT(Rf_allocVector(REALSXP, p));
  SEXP sfe = PROTECT(Rf_allocMatrix(REALSXP, p, p));
  double edof = NA_REAL;
  int rc = vc_post(h, REAL(theta), REAL(mu), REAL(sfe), Rf_asLogical(want_edof) ? &edof : NULL);
  vcb_check(rc, h);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, mu);
  SET_VECTOR_ELT(out, 1, sfe);
  SET_VECTOR_ELT(out, 2, Rf_ScalarReal(edof));
  UNPROTECT(3);
  return out;
}

/* .Call(C_vcb_predict, handle, theta, mu, newlocs, newX or NULL, newW or NULL, q, y_var) ->
 * list(eff[n_new x q], y_pred or NULL, y_var or NULL): predict.SVC_mle (R-pkg/R/predict-SVC_mle.R:80-267) */
SEXP vcb_predict(SEXP ext, SEXP theta, SEXP mu, SEXP newlocs, SEXP newX, SEXP newW, SEXP q_, SEXP want_var) {
  vc_handle* h = vcb_get(ext);
  const int64_t n_new = Rf_nrows(newlocs);
  const int q = Rf_asInteger(q_);
  const int have_xw = !Rf_isNull(newX) && !Rf_isNull(newW);
  const int var = have_xw && Rf_asLogical(want_var);
  SEXP eff = PROTECT(Rf_allocMatrix(REALSXP, (int) n_new, q));
  SEXP yp = PROTECT(have_xw ? Rf_allocVector(REALSXP, n_new) : R_NilValue);
  SEXP yv = PROTECT(var ? Rf_allocVector(REALSXP, n_new) : R_NilValue);
  int rc = vc_predict(h, REAL(theta), REAL(mu), n_new, REAL(newlocs), have_xw ? REAL(newX) : NULL,
                      have_xw ? REAL(newW) : NULL, REAL(eff), have_xw ? REAL(yp) : NULL, var ? REAL(yv) : NULL);
  vcb_check(rc, h);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, eff);
  SET_VECTOR_ELT(out, 1, yp);
  SET_VECTOR_ELT(out, 2, yv);
  UNPROTECT(4);
  return out;
}

/* .Call(C_vcb_whitened, handle, n, p) -> solve(t(R), cbind(X, y)) of the last evaluation: what CD_mu needs
 * (R-pkg/R/SVC_selection.R:30-35) without forming solve(t(R)) */
SEXP vcb_whitened(SEXP ext, SEXP n_, SEXP p_) {
  vc_handle* h = vcb_get(ext);
  SEXP Z = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(n_), Rf_asInteger(p_) + 1));
  vcb_check(vc_get_whitened(h, REAL(Z)), h);
  UNPROTECT(1);
  return Z;
}

/* .Call(C_vcb_gls_chol, R, X, y, device) -> mu as a p x 1 matrix, as GLS_chol.matrix returns it
 * (R-pkg/R/utils.R:93-101; n2LL uses it in X %*% mu) */
SEXP vcb_gls_chol(SEXP R, SEXP X, SEXP y, SEXP device) {
  const int p = Rf_ncols(X);
  SEXP mu = PROTECT(Rf_allocMatrix(REALSXP, p, 1));
  int rc = vc_gls_chol(Rf_asInteger(device), (int64_t) Rf_nrows(X), p, REAL(R), REAL(X), REAL(y), REAL(mu));
  if (rc == VC_ERR_INVALID) Rf_error("GLS_chol: R must be an upper Cholesky factor with a positive diagonal");
  if (rc != VC_OK) Rf_error("varycoef_b200 (%d): %s", rc, vc_last_error(NULL));
  UNPROTECT(1);
  return mu;
}

static const R_CallMethodDef vcb_calls[] = {
  {"C_vcb_create", (DL_FUNC) &vcb_create, 10},
  {"C_vcb_create_multi", (DL_FUNC) &vcb_create_multi, 10},
  {"C_vcb_median_dist", (DL_FUNC) &vcb_median_dist, 4},
  {"C_vcb_n2ll", (DL_FUNC) &vcb_n2ll, 4},
  {"C_vcb_n2ll_batch", (DL_FUNC) &vcb_n2ll_batch, 3},
  {"C_vcb_post", (DL_FUNC) &vcb_post, 4},
  {"C_vcb_predict", (DL_FUNC) &vcb_predict, 8},
  {"C_vcb_whitened", (DL_FUNC) &vcb_whitened, 3},
  {"C_vcb_gls_chol", (DL_FUNC) &vcb_gls_chol, 4},
  {NULL, NULL, 0}
};

void R_init_varycoef(DllInfo* dll) {
  R_registerRoutines(dll, NULL, vcb_calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
