// vc_handle.cuh -- the opaque handle behind the C ABI (shared by vc_api.cu, vc_extra.cu, vc_dist.cu).
#pragma once
#include "vc_common.cuh"

struct vc_handle {
  vc_config cfg;
  int64_t n = 0, n_pad = 0;
  int d = 0, p = 0, q = 0;
  int device = 0;
  VcMatrix M{};
  std::vector<VcColBlock> cb_host;   // block-column storage table of M (one table for all evaluation slots); device copy in cb_dev
  VcColBlock* cb_dev = nullptr;
  // O(n) inputs, zero-padded to n_pad, column-major
  double *locs = nullptr, *x = nullptr, *w = nullptr, *y = nullptr;
  double* mu_in = nullptr;           // staging for caller-supplied mu (p x slots)
  double* results = nullptr;         // device result records
  double* results_host = nullptr;    // pinned mirror
  int result_slots = 0;
  VcCholWork cw{};
  VcFinalWork fw{};
  cudaStream_t st_main = nullptr, st_panel = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // t0, after build, after chol, after finalize
  vc_timing last_timing{};
  int last_info = 0;
  int last_rec = 0;                  // result record of the last valid evaluation of the last batch
  bool factor_valid = false;
  std::vector<double> factor_theta;  // theta of the factor resident in M (empty = none)
  double* gram_dev = nullptr;        // (p+1)^2 Gram of the whitened [X y] of the last evaluation
  // Evaluation slots: M.a, M.b, cw.linv, cw.logdet_part, cw.info hold `slots` evaluations back to back.  A batch
  // (vc_n2ll_batch: optim's gradient / Hessian stencil, the analogue of the reference's optimParallel fan-out,
  // R-pkg/R/SVC_mle.R:164-200) is swept by ONE sequence of launches with the slot index in the grid, so that
  // small and mid-size problems fill the 148 SMs; slot 0 always receives the last evaluation of a batch.
  int slots = 1;
  int slot_cap_cached = 0;           // slot_cap(): queried once (cudaMemGetInfo is slow)
  VcTheta* thetas_dev = nullptr;     // slots entries: parameters of the evaluations of the group in flight
  VcTheta* thetas_host = nullptr;    // pinned staging
  int* job_slot_dev = nullptr;       // result_slots entries: slot each finalize job reads
  int* job_slot_host = nullptr;      // pinned staging
  double* mu_host = nullptr;         // pinned staging of caller-supplied mu (128 x result_slots)
  int64_t launches = 0;
  std::string err;
  // distributed state lives in vc_dist.cu
  void* dist = nullptr;
  // multi-device handle (vc_create_multi): this object owns no device memory, only one child handle per device
  std::vector<vc_handle*> kids;
  int kid_last = 0;                  // child that ran the last evaluation of the last call
};
// the handle single-target entry points act on: the child holding the last evaluation, or h itself
inline vc_handle* vc_target(vc_handle* h) { return (h && !h->kids.empty()) ? h->kids[h->kid_last] : h; }


int vc_make_theta(const vc_handle* h, const double* theta, VcTheta* t);
vc_handle* vc_alloc_common(const vc_config* cfg, int64_t n, int d, int p, int q, const double* locs,
                           const double* X, const double* W, const double* y, bool alloc_matrix);
void vc_free_handle(vc_handle* h);
int vc_fail(vc_handle* h, int code, const std::string& msg);
// computes the layout (r0 / ncols per local block column, local_rows row tiles), uploads the table and points M at it
void vc_set_layout(vc_handle* h, const std::vector<int>& r0, const std::vector<int>& ncols, int local_rows);
int vc_validate_cfg(const vc_config* cfg, int64_t n, int d, int p, int q, std::string& why);
