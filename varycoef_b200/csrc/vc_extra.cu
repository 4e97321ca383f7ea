// vc_extra.cu -- secondary entry points: GLS_chol for a caller-supplied factor and the
// next-tier rows (prep_par_output / eff_dof / predict).  Placeholder during bring-up.
#include "vc_handle.cuh"
#include <math.h>
#include <string.h>
#include <vector>

extern "C" int vc_gls_chol(int32_t, int64_t, int32_t, const double*, const double*, const double*, double*) {
  return VC_ERR_INVALID;
}
// p x p symmetric positive definite inverse (Cholesky), column-major, host
static bool spd_inverse(int p, const double* g, int ldg, double* out) {
  std::vector<double> l((size_t)p * p, 0.0);
  for (int j = 0; j < p; ++j) {
    double s = g[j + (size_t)ldg * j];
    for (int k = 0; k < j; ++k) s -= l[j + (size_t)p * k] * l[j + (size_t)p * k];
    if (!(s > 0.0)) return false;
    const double d = sqrt(s);
    l[j + (size_t)p * j] = d;
    for (int i = j + 1; i < p; ++i) {
      double t = g[i + (size_t)ldg * j];
      for (int k = 0; k < j; ++k) t -= l[i + (size_t)p * k] * l[j + (size_t)p * k];
      l[i + (size_t)p * j] = t / d;
    }
  }
  for (int c = 0; c < p; ++c) {          // solve L L^T x = e_c
    std::vector<double> w(p, 0.0);
    for (int i = 0; i < p; ++i) {
      double t = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) t -= l[i + (size_t)p * k] * w[k];
      w[i] = t / l[i + (size_t)p * i];
    }
    for (int i = p - 1; i >= 0; --i) {
      double t = w[i];
      for (int k = i + 1; k < p; ++k) t -= l[k + (size_t)p * i] * out[k + (size_t)p * c];
      out[i + (size_t)p * c] = t / l[i + (size_t)p * i];
    }
  }
  return true;
}

int vc_post_edof(vc_handle* h, const double* theta, const std::vector<double>& sigma_fe, double* edof);

extern "C" int vc_post(vc_handle* h, const double* theta, double* mu_gls, double* sigma_fe, double* edof) {
  if (!h || !theta) return VC_ERR_INVALID;
  if (h->dist) return vc_fail(h, VC_ERR_INVALID, "vc_post is not available on a distributed handle");
  const int p = h->p, m = p + 1;
  std::vector<double> mu(p);
  double val = 0.0;
  int rc = vc_n2ll(h, theta, VC_MU_GLS, nullptr, &val, mu.data(), nullptr, nullptr);
  if (rc != VC_OK) return rc;
  std::vector<double> gram((size_t)m * m), sfe((size_t)p * p);
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    VC_CUDA_TRY(cudaMemcpy(gram.data(), h->gram_dev, gram.size() * sizeof(double), cudaMemcpyDeviceToHost));
  } catch (const VcError& e) {
    return vc_fail(h, e.code, e.msg);
  }
  if (!spd_inverse(p, gram.data(), m, sfe.data()))
    return vc_fail(h, VC_ERR_NOT_PD, "X' Sigma^-1 X is not positive definite");
  if (mu_gls) memcpy(mu_gls, mu.data(), p * sizeof(double));
  if (sigma_fe) memcpy(sigma_fe, sfe.data(), sfe.size() * sizeof(double));
  if (edof) return vc_post_edof(h, theta, sfe, edof);
  return VC_OK;
}

int vc_post_edof(vc_handle* h, const double*, const std::vector<double>&, double* edof) {
  *edof = NAN;
  return vc_fail(h, VC_ERR_INVALID, "vc_post: edof (identity-border sweep) not built yet");
}
extern "C" int vc_predict(vc_handle* h, const double*, const double*, int64_t, const double*,
                          const double*, const double*, double*, double*, double*) {
  return vc_fail(h, VC_ERR_INVALID, "vc_predict: not built yet");
}
