// vc_extra.cu -- secondary entry points: GLS_chol for a caller-supplied factor and the
// next-tier rows (prep_par_output / eff_dof / predict).  Placeholder during bring-up.
#include "vc_handle.cuh"

extern "C" int vc_gls_chol(int32_t, int64_t, int32_t, const double*, const double*, const double*, double*) {
  return VC_ERR_INVALID;
}
extern "C" int vc_post(vc_handle* h, const double*, double*, double*, double*) {
  return vc_fail(h, VC_ERR_INVALID, "vc_post: not built yet");
}
extern "C" int vc_predict(vc_handle* h, const double*, const double*, int64_t, const double*,
                          const double*, const double*, double*, double*, double*) {
  return vc_fail(h, VC_ERR_INVALID, "vc_predict: not built yet");
}
