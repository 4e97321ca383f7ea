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
  const int64_t n = Rf_nrows(X);
  vc_handle* h = NULL;
  int rc = vc_create(&h, &cfg, n, Rf_ncols(locs), Rf_ncols(X), Rf_ncols(W),
                     REAL(locs), REAL(X), REAL(W), REAL(y));
  if (rc != VC_OK) Rf_error("varycoef_b200 (%d): %s", rc, vc_last_error(NULL));
  SEXP ext = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ext, vcb_finalizer, TRUE);
  UNPROTECT(1);
  return ext;
}

/* .Call(C_vcb_n2ll, handle, theta, mu or NULL) -> numeric(1) with attributes mu, logdet, quad */
SEXP vcb_n2ll(SEXP ext, SEXP theta, SEXP mu) {
  vc_handle* h = vcb_get(ext);
  const int fixed = !Rf_isNull(mu);
  double val, logdet, quad;
  SEXP mu_out = PROTECT(Rf_allocVector(REALSXP, fixed ? Rf_length(mu) : 128));
  int rc = vc_n2ll(h, REAL(theta), fixed ? VC_MU_FIXED : VC_MU_GLS, fixed ? REAL(mu) : NULL,
                   &val, REAL(mu_out), &logdet, &quad);
  vcb_check(rc, h);
  SEXP out = PROTECT(Rf_ScalarReal(val));
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

static const R_CallMethodDef vcb_calls[] = {
  {"C_vcb_create", (DL_FUNC) &vcb_create, 10},
  {"C_vcb_n2ll", (DL_FUNC) &vcb_n2ll, 3},
  {"C_vcb_n2ll_batch", (DL_FUNC) &vcb_n2ll_batch, 3},
  {NULL, NULL, 0}
};

void R_init_varycoef(DllInfo* dll) {
  R_registerRoutines(dll, NULL, vcb_calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
