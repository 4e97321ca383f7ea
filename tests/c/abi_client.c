/* Plain-C client of the C ABI (include/varycoef_b200.h): what a non-Python host links against.
 * Built and run by tests/test_host_logic.py.  Without a GPU it checks the argument validation and that a
 * valid vc_create fails loudly with VC_ERR_CUDA; with a GPU it evaluates a tiny model and prints -2logL. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "varycoef_b200.h"

int main(void) {
  vc_config cfg;
  if (vc_config_default(&cfg) != VC_OK) return 10;
  if (vc_config_default(NULL) != VC_ERR_INVALID) return 11;
  printf("abi %d cov %d dist %d taper %d\n", vc_abi_version(), cfg.cov_id, cfg.dist_method, cfg.taper_id);
  enum { N = 40, D = 1, P = 2, Q = 2 };
  double locs[N], X[N * P], y[N];
  for (int i = 0; i < N; ++i) {
    locs[i] = 0.25 * i;
    X[i] = 1.0;
    X[N + i] = sin(0.7 * i);
    y[i] = 1.0 + 2.0 * X[N + i] + 0.3 * cos(1.3 * i);
  }
  vc_handle* h = NULL;
  if (vc_create(&h, &cfg, 0, D, P, Q, locs, X, X, y) != VC_ERR_INVALID) return 12;       /* n < 1 */
  if (vc_create(&h, &cfg, N, D, P, Q, NULL, X, X, y) != VC_ERR_INVALID) return 13;       /* NULL input */
  cfg.cov_id = 99;
  if (vc_create(&h, &cfg, N, D, P, Q, locs, X, X, y) != VC_ERR_INVALID) return 14;       /* unknown covariance */
  vc_config_default(&cfg);
  int rc = vc_create(&h, &cfg, N, D, P, Q, locs, X, X, y);
  if (rc == VC_ERR_CUDA) {
    printf("no gpu: %s\n", vc_last_error(NULL));
    return 0;
  }
  if (rc != VC_OK) return 15;
  const double theta[2 * Q + 1] = {1.5, 0.8, 1.0, 0.4, 0.3};
  double val = 0, mu[P], logdet = 0, quad = 0;
  rc = vc_n2ll(h, theta, VC_MU_GLS, NULL, &val, mu, &logdet, &quad);
  if (rc != VC_OK) { printf("error: %s\n", vc_last_error(h)); vc_destroy(h); return 16; }
  printf("n2ll %.12f mu %.9f %.9f logdet %.9f quad %.9f\n", val, mu[0], mu[1], logdet, quad);
  const double bad[2 * Q + 1] = {-1.0, 0.8, 1.0, 0.4, 0.3};
  if (vc_n2ll(h, bad, VC_MU_GLS, NULL, &val, mu, NULL, NULL) != VC_ERR_INVALID) { vc_destroy(h); return 17; }
  /* a batch: three thetas, the first and the last equal (they share one factorisation), bit-identical to vc_n2ll */
  const double th3[3 * (2 * Q + 1)] = {1.5, 0.8, 1.0, 0.4, 0.3, 1.6, 0.8, 1.0, 0.4, 0.3, 1.5, 0.8, 1.0, 0.4, 0.3};
  double v3[3];
  int32_t st3[3];
  if (vc_n2ll_batch(h, 3, th3, VC_MU_GLS, NULL, v3, NULL, st3) != VC_OK) { vc_destroy(h); return 18; }
  if (v3[0] != v3[2] || st3[1] != VC_OK) { vc_destroy(h); return 19; }
  if (vc_n2ll(h, theta, VC_MU_GLS, NULL, &val, mu, NULL, NULL) != VC_OK || val != v3[0]) { vc_destroy(h); return 20; }
  /* med_dist of MLE_computation: 40 equispaced points at 0.25 -> median of all 1600 |i - j| / 4 */
  double med = -1.0;
  if (vc_median_dist(0, N, D, locs, VC_DIST_EUCLIDEAN, 2.0, &med, NULL) != VC_OK) { vc_destroy(h); return 21; }
  printf("batch %.12f %.12f median_dist %.6f devices %d\n", v3[0], v3[1], med, vc_device_count(h));
  vc_destroy(h);
  return 0;
}
