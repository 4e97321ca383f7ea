// vc_api.cu -- the C ABI (include/varycoef_b200.h): handle management and the per-evaluation
// orchestration  build (K1) -> bordered Cholesky (K2) -> finalize (K3)  on CUDA streams.
#include "vc_handle.cuh"
#include <math.h>
#include <string.h>
#include <limits.h>
#include <new>
#include <algorithm>

static thread_local std::string g_create_error;

extern int vc_dist_n2ll(vc_handle* h, const double* theta, int32_t mu_mode, const double* mu_in,
                        double* n2ll, double* mu_out, double* logdet, double* quad);
extern void vc_dist_destroy(vc_handle* h);

int vc_fail(vc_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

extern "C" int vc_abi_version(void) { return VC_ABI_VERSION; }

extern "C" int vc_config_default(vc_config* cfg) {
  if (!cfg) return VC_ERR_INVALID;
  memset(cfg, 0, sizeof(*cfg));
  cfg->cov_id = VC_COV_EXP;
  cfg->dist_method = VC_DIST_EUCLIDEAN;
  cfg->minkowski_p = 2.0;
  cfg->taper_id = VC_TAPER_NONE;
  cfg->taper_range = 0.0;
  cfg->device = 0;
  cfg->nb_outer = 0;
  cfg->flags = 0;
  return VC_OK;
}

static void upload_padded(double* dst, const double* src, int64_t n, int64_t n_pad, int cols,
                          cudaStream_t st) {
  // column-major n x cols -> n_pad x cols (tail rows stay zero)
  VC_CUDA_TRY(cudaMemcpy2DAsync(dst, n_pad * sizeof(double), src, n * sizeof(double),
                                n * sizeof(double), cols, cudaMemcpyHostToDevice, st));
}

int vc_validate_cfg(const vc_config* cfg, int64_t n, int d, int p, int q, std::string& why) {
  if (n < 1) { why = "n must be >= 1"; return VC_ERR_INVALID; }
  if (d < 1 || d > VC_MAXD) { why = "d (coordinate columns) must be in 1..8"; return VC_ERR_INVALID; }
  if (p < 1 || p + 1 > VC_BORDER_ROWS) { why = "p must be in 1..127"; return VC_ERR_INVALID; }
  if (q < 1 || q > VC_MAXQ) { why = "q must be in 1..32"; return VC_ERR_INVALID; }
  if (cfg->cov_id < 0 || cfg->cov_id > VC_COV_WEND2) { why = "unknown cov_id"; return VC_ERR_INVALID; }
  if (cfg->dist_method < 0 || cfg->dist_method > VC_DIST_MINKOWSKI) { why = "unknown dist_method"; return VC_ERR_INVALID; }
  if (cfg->taper_id < 0 || cfg->taper_id > VC_TAPER_WEND2) { why = "unknown taper_id"; return VC_ERR_INVALID; }
  if (cfg->taper_id != VC_TAPER_NONE && !(cfg->taper_range > 0.0)) { why = "taper_range must be > 0"; return VC_ERR_INVALID; }
  if (cfg->dist_method == VC_DIST_MINKOWSKI && !(cfg->minkowski_p > 0.0)) { why = "minkowski_p must be > 0"; return VC_ERR_INVALID; }
  if (cfg->nb_outer < 0 || cfg->nb_outer % VC_TILE != 0) { why = "nb_outer must be a multiple of 128"; return VC_ERR_INVALID; }
  return VC_OK;
}

static void free_lane(VcLane* L);

void vc_free_handle(vc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->dist) vc_dist_destroy(h);
  for (VcLane* L : h->lanes) free_lane(L);
  h->lanes.clear();
  if (h->ev_inputs) cudaEventDestroy(h->ev_inputs);
  cudaFree(h->cb_dev);
  cudaFree(h->M.a); cudaFree(h->M.b); vc_chol_free_plans(h->cw); cudaFree(h->locs); cudaFree(h->x); cudaFree(h->w); cudaFree(h->y);
  cudaFree(h->mu_in); cudaFree(h->results); cudaFree(h->gram_dev);
  if (h->results_host) cudaFreeHost(h->results_host);
  cudaFree(h->cw.linv); cudaFree(h->cw.logdet_part); cudaFree(h->cw.info);
  cudaFree(h->fw.gram_part); cudaFree(h->fw.quad_part);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->cw.ev_panel) cudaEventDestroy(h->cw.ev_panel);
  if (h->cw.ev_update) cudaEventDestroy(h->cw.ev_update);
  if (h->cw.ev_col) cudaEventDestroy(h->cw.ev_col);
  if (h->st_main) cudaStreamDestroy(h->st_main);
  if (h->st_panel) cudaStreamDestroy(h->st_panel);
  delete h;
}

static void ensure_result_slots(vc_handle* h, int m) {
  if (m <= h->result_slots) return;
  cudaFree(h->results); cudaFree(h->mu_in);
  if (h->results_host) cudaFreeHost(h->results_host);
  h->results = nullptr; h->mu_in = nullptr; h->results_host = nullptr; h->result_slots = 0;
  VC_CUDA_TRY(cudaMalloc(&h->results, (size_t)m * VC_RES_STRIDE * sizeof(double)));
  VC_CUDA_TRY(cudaMalloc(&h->mu_in, (size_t)m * 128 * sizeof(double)));
  VC_CUDA_TRY(cudaMallocHost(&h->results_host, (size_t)m * VC_RES_STRIDE * sizeof(double)));
  h->result_slots = m;
}

void vc_set_layout(vc_handle* h, const std::vector<int>& r0, const std::vector<int>& ncols, int local_rows) {
  h->M.a_doubles = vc_layout_fill(h->cb_host, r0, ncols, local_rows);
  h->M.blk = h->cw.nb_tiles;
  cudaFree(h->cb_dev);
  h->cb_dev = nullptr;
  VC_CUDA_TRY(cudaMalloc(&h->cb_dev, std::max<size_t>(h->cb_host.size(), 1) * sizeof(VcColBlock)));
  if (!h->cb_host.empty())
    VC_CUDA_TRY(cudaMemcpy(h->cb_dev, h->cb_host.data(), h->cb_host.size() * sizeof(VcColBlock), cudaMemcpyHostToDevice));
  h->M.cb = h->cb_dev;
  h->M.cb_host = h->cb_host.data();
}

// shared by vc_create and vc_create_dist (which passes its own local matrix shape later)
vc_handle* vc_alloc_common(const vc_config* cfg, int64_t n, int d, int p, int q, const double* locs,
                           const double* X, const double* W, const double* y, bool alloc_matrix) {
  vc_handle* h = new vc_handle();
  h->cfg = *cfg;
  h->n = n; h->d = d; h->p = p; h->q = q; h->device = cfg->device;
  h->n_pad = (n + VC_TILE - 1) / VC_TILE * VC_TILE;
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    int lo = 0, hi = 0;
    VC_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    VC_CUDA_TRY(cudaStreamCreateWithPriority(&h->st_main, cudaStreamNonBlocking, lo));
    VC_CUDA_TRY(cudaStreamCreateWithPriority(&h->st_panel, cudaStreamNonBlocking, hi));
    for (auto& e : h->ev) VC_CUDA_TRY(cudaEventCreate(&e));
    VC_CUDA_TRY(cudaEventCreateWithFlags(&h->cw.ev_panel, cudaEventDisableTiming));
    VC_CUDA_TRY(cudaEventCreateWithFlags(&h->cw.ev_update, cudaEventDisableTiming));
    VC_CUDA_TRY(cudaEventCreateWithFlags(&h->cw.ev_col, cudaEventDisableTiming));
    const int64_t np = h->n_pad;
    VC_CUDA_TRY(cudaMalloc(&h->locs, np * d * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&h->x, np * p * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&h->w, np * q * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&h->y, np * sizeof(double)));
    VC_CUDA_TRY(cudaMemsetAsync(h->locs, 0, np * d * sizeof(double), h->st_main));
    VC_CUDA_TRY(cudaMemsetAsync(h->x, 0, np * p * sizeof(double), h->st_main));
    VC_CUDA_TRY(cudaMemsetAsync(h->w, 0, np * q * sizeof(double), h->st_main));
    VC_CUDA_TRY(cudaMemsetAsync(h->y, 0, np * sizeof(double), h->st_main));
    upload_padded(h->locs, locs, n, np, d, h->st_main);
    upload_padded(h->x, X, n, np, p, h->st_main);
    upload_padded(h->w, W, n, np, q, h->st_main);
    upload_padded(h->y, y, n, np, 1, h->st_main);
    ensure_result_slots(h, 64);
    h->M.n = n; h->M.n_pad = np; h->M.nt = (int)(np / VC_TILE);
    h->M.nbt = 1;
    h->M.ldb = VC_BORDER_ROWS;
    h->cw.legacy_syrk = (cfg->flags & VC_FLAG_LEGACY_SYRK) != 0;
    h->cw.legacy_potrf = (cfg->flags & VC_FLAG_LEGACY_POTRF) != 0;
    {
      cudaDeviceProp prop;
      VC_CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
      h->cw.num_sms = prop.multiProcessorCount;
    }
    // outer block: measured best 256 for n <= 12k, 512 up to 40k, 1024 beyond (profiles/r01_block_sweep.md)
    // measured (profiles/r01_block_sweep.md, r01_panel_chain.md): 128 up to n = 4096, 256 up to 12288, 512, 1024 from 40000
    const int nb_auto = np >= 40000 ? 1024 : (np <= 4096 ? 128 : (np <= 12288 ? 256 : 512));
    h->cw.nb_tiles = (cfg->nb_outer > 0 ? cfg->nb_outer : nb_auto) / VC_TILE;
    h->cw.lookahead = !(cfg->flags & VC_FLAG_NO_LOOKAHEAD);
    h->cw.st_main = h->st_main; h->cw.st_panel = h->st_panel;
    if (alloc_matrix) {
      // lower block-column trapezoids only: block column c starts at row tile c * blk
      const int blk = h->cw.nb_tiles, nt = h->M.nt;
      std::vector<int> r0, ncols;
      for (int c = 0; c * blk < nt; ++c) { r0.push_back(c * blk); ncols.push_back(std::min(blk, nt - c * blk)); }
      vc_set_layout(h, r0, ncols, nt);
      VC_CUDA_TRY(cudaMalloc(&h->M.a, (size_t)h->M.a_doubles * sizeof(double)));
      VC_CUDA_TRY(cudaMalloc(&h->M.b, (size_t)h->M.ldb * np * sizeof(double)));
      VC_CUDA_TRY(cudaMalloc(&h->cw.linv, (size_t)h->M.nt * VC_TILE * VC_TILE * sizeof(double)));
      // the diagonal-tile kernel writes only the lower 16x16 blocks of each inverse; the rest stays zero
      // (on the handle's own stream: a legacy-stream memset is not ordered against non-blocking streams)
      VC_CUDA_TRY(cudaMemsetAsync(h->cw.linv, 0, (size_t)h->M.nt * VC_TILE * VC_TILE * sizeof(double), h->st_main));
      VC_CUDA_TRY(cudaMalloc(&h->cw.logdet_part, (size_t)h->M.nt * sizeof(double)));
    }
    VC_CUDA_TRY(cudaMalloc(&h->cw.info, sizeof(int)));
    int blocks = (int)(n / 512);
    if (blocks < 1) blocks = 1;
    if (blocks > 148) blocks = 148;
    h->fw.blocks = blocks;
    const int m = p + 1;
    VC_CUDA_TRY(cudaMalloc(&h->fw.gram_part, (size_t)blocks * m * (m + 1) * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&h->fw.quad_part, (size_t)blocks * 2 * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&h->gram_dev, (size_t)m * m * sizeof(double)));
    vc_build_init();
    vc_chol_init();
    VC_CUDA_TRY(cudaStreamSynchronize(h->st_main));
  } catch (const VcError& e) {
    g_create_error = e.msg;
    vc_free_handle(h);
    throw;
  }
  return h;
}

extern "C" int vc_create(vc_handle** out, const vc_config* cfg_in, int64_t n, int32_t d, int32_t p,
                         int32_t q, const double* locs, const double* X, const double* W,
                         const double* y) {
  if (!out) return VC_ERR_INVALID;
  *out = nullptr;
  vc_config cfg;
  if (cfg_in) cfg = *cfg_in; else vc_config_default(&cfg);
  if (!locs || !X || !W || !y) return vc_fail(nullptr, VC_ERR_INVALID, "NULL input pointer");
  std::string why;
  if (vc_validate_cfg(&cfg, n, d, p, q, why) != VC_OK) return vc_fail(nullptr, VC_ERR_INVALID, why);
  try {
    *out = vc_alloc_common(&cfg, n, d, p, q, locs, X, W, y, true);
  } catch (const VcError& e) {
    return vc_fail(nullptr, e.code, e.msg);
  } catch (const std::bad_alloc&) {
    return vc_fail(nullptr, VC_ERR_OOM, "host allocation failed");
  }
  return VC_OK;
}

extern "C" void vc_destroy(vc_handle* h) { vc_free_handle(h); }

extern "C" int vc_set_data(vc_handle* h, const double* locs, const double* X, const double* W,
                           const double* y) {
  if (!h) return VC_ERR_INVALID;
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    if (locs) upload_padded(h->locs, locs, h->n, h->n_pad, h->d, h->st_main);
    if (X) upload_padded(h->x, X, h->n, h->n_pad, h->p, h->st_main);
    if (W) upload_padded(h->w, W, h->n, h->n_pad, h->q, h->st_main);
    if (y) upload_padded(h->y, y, h->n, h->n_pad, 1, h->st_main);
    h->factor_valid = false;
    h->factor_theta.clear();
    for (VcLane* L : h->lanes) L->factor_theta.clear();
  } catch (const VcError& e) {
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}

// theta -> kernel parameter block; VC_ERR_INVALID when a value cannot define a covariance
int vc_make_theta(const vc_handle* h, const double* theta, VcTheta* t) {
  memset(t, 0, sizeof(*t));
  t->q = h->q;
  t->cov_id = h->cfg.cov_id;
  t->taper_id = h->cfg.taper_id;
  t->dist_method = h->cfg.dist_method;
  t->mink_p = h->cfg.minkowski_p;
  t->inv_delta = h->cfg.taper_id != VC_TAPER_NONE ? 1.0 / h->cfg.taper_range : 0.0;
  for (int k = 0; k < h->q; ++k) {
    const double rho = theta[2 * k], s2 = theta[2 * k + 1];
    if (!(rho > 0.0) || !isfinite(rho) || !isfinite(s2)) return VC_ERR_INVALID;
    t->inv_range[k] = 1.0 / rho;
    t->sill[k] = s2;
  }
  t->tau2 = theta[2 * h->q];
  if (!isfinite(t->tau2)) return VC_ERR_INVALID;
  return VC_OK;
}

// how many evaluation lanes a batch may use: one evaluation fills the GPU from n ~ 20k on
static int lane_cap(const vc_handle* h) {
  if (h->dist) return 1;
  static const int env_cap = getenv("VC_MAX_LANES") ? atoi(getenv("VC_MAX_LANES")) : 0;   // tuning / debugging
  if (env_cap >= 1) return std::min(env_cap, 4);
  if (h->n_pad <= 12288) return 4;
  if (h->n_pad <= 24576) return 2;
  return 1;
}

static void free_lane(VcLane* L) {
  if (!L) return;
  cudaFree(L->M.a); cudaFree(L->M.b); vc_chol_free_plans(L->cw);
  cudaFree(L->cw.linv); cudaFree(L->cw.logdet_part); cudaFree(L->cw.info);
  cudaFree(L->fw.gram_part); cudaFree(L->fw.quad_part); cudaFree(L->gram_dev);
  if (L->cw.ev_panel) cudaEventDestroy(L->cw.ev_panel);
  if (L->cw.ev_update) cudaEventDestroy(L->cw.ev_update);
  if (L->cw.ev_col) cudaEventDestroy(L->cw.ev_col);
  if (L->ev_done) cudaEventDestroy(L->ev_done);
  if (L->st_main) cudaStreamDestroy(L->st_main);
  if (L->st_panel) cudaStreamDestroy(L->st_panel);
  delete L;
}

// A lane joins h->lanes only when every stream, event and buffer of it exists: a half-built lane must
// never be visible to a later batch (it would launch on null pointers).  Throws VcError on failure.
static void make_lane(vc_handle* h) {
  VcLane* L = new VcLane();
  try {
    const int64_t np = h->n_pad;
    int lo = 0, hi = 0;
    VC_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    VC_CUDA_TRY(cudaStreamCreateWithPriority(&L->st_main, cudaStreamNonBlocking, lo));
    VC_CUDA_TRY(cudaStreamCreateWithPriority(&L->st_panel, cudaStreamNonBlocking, hi));
    VC_CUDA_TRY(cudaEventCreateWithFlags(&L->cw.ev_panel, cudaEventDisableTiming));
    VC_CUDA_TRY(cudaEventCreateWithFlags(&L->cw.ev_update, cudaEventDisableTiming));
    VC_CUDA_TRY(cudaEventCreateWithFlags(&L->cw.ev_col, cudaEventDisableTiming));
    VC_CUDA_TRY(cudaEventCreateWithFlags(&L->ev_done, cudaEventDisableTiming));
    L->M = h->M;
    L->M.a = nullptr; L->M.b = nullptr;
    L->M.nbt = 1; L->M.ldb = VC_BORDER_ROWS;
    VC_CUDA_TRY(cudaMalloc(&L->M.a, (size_t)L->M.a_doubles * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&L->M.b, (size_t)L->M.ldb * np * sizeof(double)));
    L->cw.nb_tiles = h->cw.nb_tiles; L->cw.lookahead = h->cw.lookahead;
    L->cw.legacy_syrk = h->cw.legacy_syrk; L->cw.legacy_potrf = h->cw.legacy_potrf; L->cw.num_sms = h->cw.num_sms;
    L->cw.st_main = L->st_main; L->cw.st_panel = L->st_panel;
    VC_CUDA_TRY(cudaMalloc(&L->cw.linv, (size_t)L->M.nt * VC_TILE * VC_TILE * sizeof(double)));
    VC_CUDA_TRY(cudaMemsetAsync(L->cw.linv, 0, (size_t)L->M.nt * VC_TILE * VC_TILE * sizeof(double), L->st_main));
    VC_CUDA_TRY(cudaMalloc(&L->cw.logdet_part, (size_t)L->M.nt * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&L->cw.info, sizeof(int)));
    const int m = h->p + 1;
    L->fw.blocks = h->fw.blocks;
    VC_CUDA_TRY(cudaMalloc(&L->fw.gram_part, (size_t)L->fw.blocks * m * (m + 1) * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&L->fw.quad_part, (size_t)L->fw.blocks * 2 * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&L->gram_dev, (size_t)m * m * sizeof(double)));
  } catch (...) {
    free_lane(L);
    throw;
  }
  h->lanes.push_back(L);
}

// enqueue one evaluation on a lane (lane < 0: the handle's own matrix and streams); result record
// `slot`.  When theta equals the theta of the factor already resident in that lane (optim's probes of
// the mean parameters in the non-profile objective, vc_post after a fit), only the O(n p) finalize runs.
static void enqueue_eval(vc_handle* h, int lane, const VcTheta& th, const double* theta, int mu_mode,
                         int slot, bool timed) {
  VcLane* L = lane >= 0 ? h->lanes[lane] : nullptr;
  VcMatrix& M = L ? L->M : h->M;
  VcCholWork& cw = L ? L->cw : h->cw;
  VcFinalWork& fw = L ? L->fw : h->fw;
  std::vector<double>& ft = L ? L->factor_theta : h->factor_theta;
  cudaStream_t st = L ? L->st_main : h->st_main;
  const size_t nth = (size_t)2 * h->q + 1;
  const bool cached = ft.size() == nth && memcmp(ft.data(), theta, nth * sizeof(double)) == 0;
  if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[0], st));
  if (!cached) {
    vc_launch_build(M, th, h->d, h->locs, h->w, st, &h->launches);
    vc_launch_border(M, h->p, h->x, h->y, st, &h->launches);
  }
  if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[1], st));
  if (!cached) {
    M.nbt = 1;
    vc_chol_factor(M, cw, &h->launches);
    ft.assign(theta, theta + nth);
  }
  if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[2], st));
  vc_launch_finalize(M, h->p, mu_mode, h->mu_in + (size_t)slot * 128, cw.logdet_part, cw.info, fw,
                     h->results + (size_t)slot * VC_RES_STRIDE, st, &h->launches, L ? L->gram_dev : h->gram_dev);
  if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[3], st));
}

extern "C" int vc_n2ll_batch(vc_handle* h, int32_t m, const double* thetas, int32_t mu_mode,
                             const double* mus_in, double* out_n2ll, double* out_mu,
                             int32_t* out_status) {
  if (!h) return VC_ERR_INVALID;
  if (m < 1 || !thetas || !out_n2ll) return vc_fail(h, VC_ERR_INVALID, "bad batch arguments");
  if (mu_mode != VC_MU_GLS && mu_mode != VC_MU_FIXED) return vc_fail(h, VC_ERR_INVALID, "unknown mu_mode");
  if (mu_mode == VC_MU_FIXED && !mus_in) return vc_fail(h, VC_ERR_INVALID, "mu_mode FIXED needs mu");
  if (h->dist) {
    int first = VC_OK;
    for (int e = 0; e < m; ++e) {
      int s = vc_dist_n2ll(h, thetas + (size_t)e * (2 * h->q + 1), mu_mode,
                           mus_in ? mus_in + (size_t)e * h->p : nullptr, out_n2ll + e,
                           out_mu ? out_mu + (size_t)e * h->p : nullptr, nullptr, nullptr);
      if (out_status) out_status[e] = s;
      if (s != VC_OK && first == VC_OK) first = s;
    }
    return first;
  }
  const int nth = 2 * h->q + 1;
  int first = VC_OK;
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    ensure_result_slots(h, m);
    std::vector<int> status(m, VC_OK);
    if (mu_mode == VC_MU_FIXED)
      VC_CUDA_TRY(cudaMemcpy2DAsync(h->mu_in, 128 * sizeof(double), mus_in, h->p * sizeof(double),
                                    h->p * sizeof(double), m, cudaMemcpyHostToDevice, h->st_main));
    // lanes: evaluation e runs on lane (m - 1 - e) % nl, so the LAST one lands on the handle's own
    // matrix (vc_get_chol / vc_get_whitened / vc_post refer to it)
    int nl = std::min(m, lane_cap(h));
    while ((int)h->lanes.size() < nl - 1) {
      try {
        make_lane(h);
      } catch (const VcError& e) {
        if (e.code != VC_ERR_OOM) throw;
        (void)cudaGetLastError();                 // out of device memory: run the batch on the lanes that exist
        nl = (int)h->lanes.size() + 1;
      }
    }
    if (nl > 1) {
      if (!h->ev_inputs) VC_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_inputs, cudaEventDisableTiming));
      VC_CUDA_TRY(cudaEventRecord(h->ev_inputs, h->st_main));   // inputs / mu staged on st_main
      for (int l = 0; l < nl - 1; ++l) VC_CUDA_TRY(cudaStreamWaitEvent(h->lanes[l]->st_main, h->ev_inputs, 0));
    }
    for (int e = 0; e < m; ++e) {
      VcTheta th;
      if (vc_make_theta(h, thetas + (size_t)e * nth, &th) != VC_OK) {
        status[e] = VC_ERR_INVALID;
        continue;
      }
      const int lane = (m - 1 - e) % nl;              // 0 = the handle itself
      enqueue_eval(h, lane - 1, th, thetas + (size_t)e * nth, mu_mode, e, e == m - 1);
    }
    for (int l = 0; l < nl - 1; ++l) {
      VC_CUDA_TRY(cudaEventRecord(h->lanes[l]->ev_done, h->lanes[l]->st_main));
      VC_CUDA_TRY(cudaStreamWaitEvent(h->st_main, h->lanes[l]->ev_done, 0));
    }
    VC_CUDA_TRY(cudaMemcpyAsync(h->results_host, h->results, (size_t)m * VC_RES_STRIDE * sizeof(double),
                                cudaMemcpyDeviceToHost, h->st_main));
    VC_CUDA_TRY(cudaStreamSynchronize(h->st_main));
    VC_CUDA_TRY(cudaGetLastError());
    for (int e = 0; e < m; ++e) {
      const double* r = h->results_host + (size_t)e * VC_RES_STRIDE;
      if (status[e] == VC_OK) {
        const int info = (int)r[VC_RES_INFO];
        if (info != 0) { status[e] = VC_ERR_NOT_PD; h->last_info = info; }
        out_n2ll[e] = r[VC_RES_N2LL];
        if (out_mu) for (int i = 0; i < h->p; ++i) out_mu[(size_t)e * h->p + i] = r[VC_RES_MU + i];
      } else {
        out_n2ll[e] = NAN;
        if (out_mu) for (int i = 0; i < h->p; ++i) out_mu[(size_t)e * h->p + i] = NAN;
      }
      if (out_status) out_status[e] = status[e];
      if (status[e] != VC_OK && first == VC_OK) first = status[e];
    }
    if (status[m - 1] == VC_OK || status[m - 1] == VC_ERR_NOT_PD) {
      float b = 0, c = 0, f = 0;
      cudaEventElapsedTime(&b, h->ev[0], h->ev[1]);
      cudaEventElapsedTime(&c, h->ev[1], h->ev[2]);
      cudaEventElapsedTime(&f, h->ev[2], h->ev[3]);
      h->last_timing = {b, c, f, b + c + f};
    }
    h->factor_valid = (status[m - 1] == VC_OK);
    if (first == VC_ERR_NOT_PD) {
      char buf[128];
      snprintf(buf, sizeof(buf), "the leading minor of order %d is not positive", h->last_info);
      h->err = buf;
    } else if (first == VC_ERR_INVALID) {
      h->err = "theta does not define a covariance (range <= 0 or non-finite value)";
    }
  } catch (const VcError& e) {
    h->factor_theta.clear();
    for (VcLane* L : h->lanes) L->factor_theta.clear();
    h->factor_valid = false;
    return vc_fail(h, e.code, e.msg);
  }
  return first;
}

extern "C" int vc_n2ll(vc_handle* h, const double* theta, int32_t mu_mode, const double* mu_in,
                       double* n2ll, double* mu_out, double* logdet, double* quad) {
  if (!h || !theta || !n2ll) return VC_ERR_INVALID;
  if (h->dist) return vc_dist_n2ll(h, theta, mu_mode, mu_in, n2ll, mu_out, logdet, quad);
  int32_t status = VC_OK;
  h->last_info = 0;
  int rc = vc_n2ll_batch(h, 1, theta, mu_mode, mu_in, n2ll, mu_out, &status);
  if (rc == VC_OK || rc == VC_ERR_NOT_PD) {
    if (logdet) *logdet = h->results_host[VC_RES_LOGDET];
    if (quad) *quad = h->results_host[VC_RES_QUAD];
  }
  return rc;
}

extern "C" int vc_sigma(vc_handle* h, const double* theta, double* out) {
  if (!h || !theta || !out) return VC_ERR_INVALID;
  if (h->dist) return vc_fail(h, VC_ERR_INVALID, "vc_sigma is not available on a distributed handle");
  VcTheta th;
  if (vc_make_theta(h, theta, &th) != VC_OK) return vc_fail(h, VC_ERR_INVALID, "invalid theta");
  double* dev = nullptr;
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    VC_CUDA_TRY(cudaMalloc(&dev, (size_t)h->n * h->n * sizeof(double)));
    vc_launch_sigma_full(dev, h->n, h->n_pad, th, h->d, h->locs, h->w, h->st_main, &h->launches);
    VC_CUDA_TRY(cudaMemcpyAsync(out, dev, (size_t)h->n * h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st_main));
    VC_CUDA_TRY(cudaStreamSynchronize(h->st_main));
    cudaFree(dev);
  } catch (const VcError& e) {
    cudaFree(dev);
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}

extern "C" int vc_sigma_device(vc_handle* h, const double* theta, double* out_dev) {
  if (!h || !theta || !out_dev) return VC_ERR_INVALID;
  if (h->dist) return vc_fail(h, VC_ERR_INVALID, "vc_sigma_device is not available on a distributed handle");
  VcTheta th;
  if (vc_make_theta(h, theta, &th) != VC_OK) return vc_fail(h, VC_ERR_INVALID, "invalid theta");
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    cudaPointerAttributes attr;
    VC_CUDA_TRY(cudaPointerGetAttributes(&attr, out_dev));
    if (attr.type != cudaMemoryTypeDevice || attr.device != h->device)
      return vc_fail(h, VC_ERR_INVALID, "out_dev must be device memory of the handle's GPU");
    vc_launch_sigma_full(out_dev, h->n, h->n_pad, th, h->d, h->locs, h->w, h->st_main, &h->launches);
    VC_CUDA_TRY(cudaStreamSynchronize(h->st_main));
  } catch (const VcError& e) {
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}

extern "C" int vc_get_chol(vc_handle* h, double* R_out) {
  if (!h || !R_out) return VC_ERR_INVALID;
  if (h->dist) return vc_fail(h, VC_ERR_INVALID, "vc_get_chol is not available on a distributed handle");
  if (!h->factor_valid) return vc_fail(h, VC_ERR_INVALID, "no valid factor: call vc_n2ll first");
  double* dev = nullptr;
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    VC_CUDA_TRY(cudaMalloc(&dev, (size_t)h->n * h->n * sizeof(double)));
    vc_launch_extract_chol(h->M, dev, h->st_main, &h->launches);
    VC_CUDA_TRY(cudaMemcpyAsync(R_out, dev, (size_t)h->n * h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st_main));
    VC_CUDA_TRY(cudaStreamSynchronize(h->st_main));
    cudaFree(dev);
  } catch (const VcError& e) {
    cudaFree(dev);
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}

extern "C" int vc_get_whitened(vc_handle* h, double* Z_out) {
  if (!h || !Z_out) return VC_ERR_INVALID;
  if (h->dist) return vc_fail(h, VC_ERR_INVALID, "vc_get_whitened is not available on a distributed handle");
  if (!h->factor_valid) return vc_fail(h, VC_ERR_INVALID, "no valid factor: call vc_n2ll first");
  double* dev = nullptr;
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->n * (h->p + 1) * sizeof(double);
    VC_CUDA_TRY(cudaMalloc(&dev, bytes));
    vc_launch_gather_whitened(h->M, h->p, dev, h->st_main, &h->launches);
    VC_CUDA_TRY(cudaMemcpyAsync(Z_out, dev, bytes, cudaMemcpyDeviceToHost, h->st_main));
    VC_CUDA_TRY(cudaStreamSynchronize(h->st_main));
    cudaFree(dev);
  } catch (const VcError& e) {
    cudaFree(dev);
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}

extern "C" int vc_timings(vc_handle* h, vc_timing* out) {
  if (!h || !out) return VC_ERR_INVALID;
  *out = h->last_timing;
  return VC_OK;
}
extern "C" int vc_last_info(vc_handle* h) { return h ? h->last_info : 0; }
extern "C" const char* vc_last_error(vc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" int64_t vc_padded_n(vc_handle* h) { return h ? h->n_pad : 0; }
extern "C" int64_t vc_launch_count(vc_handle* h) { return h ? h->launches : 0; }
