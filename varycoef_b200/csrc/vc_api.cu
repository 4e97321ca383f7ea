// vc_api.cu -- the C ABI (include/varycoef_b200.h): handle management and the per-evaluation
// orchestration  build (K1) -> bordered Cholesky (K2) -> finalize (K3)  on CUDA streams.
#include "vc_handle.cuh"
#include <math.h>
#include <string.h>
#include <limits.h>
#include <new>
#include <algorithm>
#include <thread>

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

// outer Cholesky block = storage block.  Measured with batches of 16 (the stencils optim evaluates) and single
// evaluations (profiles/r02_batch_small_n.md): 256 up to n = 4 096, 512 up to 40 000, 1024 beyond
static int auto_outer_block(int64_t n_pad) { return n_pad >= 40000 ? 1024 : (n_pad <= 4096 ? 256 : 512); }

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

void vc_free_handle(vc_handle* h) {
  if (!h) return;
  if (!h->kids.empty()) {
    for (vc_handle* k : h->kids) vc_free_handle(k);
    delete h;
    return;
  }
  cudaSetDevice(h->device);
  if (h->dist) vc_dist_destroy(h);
  cudaFree(h->cb_dev);
  cudaFree(h->M.a); cudaFree(h->M.b); vc_chol_free_plans(h->cw); cudaFree(h->locs); cudaFree(h->x); cudaFree(h->w); cudaFree(h->y);
  cudaFree(h->mu_in); cudaFree(h->results); cudaFree(h->gram_dev);
  cudaFree(h->thetas_dev); cudaFree(h->job_slot_dev);
  if (h->thetas_host) cudaFreeHost(h->thetas_host);
  if (h->job_slot_host) cudaFreeHost(h->job_slot_host);
  if (h->mu_host) cudaFreeHost(h->mu_host);
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

// per-JOB buffers (one job = one requested evaluation): result records, mu, the slot it reads, finalize partials
static void ensure_result_slots(vc_handle* h, int m) {
  if (m <= h->result_slots) return;
  cudaFree(h->results); cudaFree(h->mu_in); cudaFree(h->job_slot_dev); cudaFree(h->fw.gram_part); cudaFree(h->fw.quad_part);
  if (h->results_host) cudaFreeHost(h->results_host);
  if (h->job_slot_host) cudaFreeHost(h->job_slot_host);
  if (h->mu_host) cudaFreeHost(h->mu_host);
  if (h->thetas_host) cudaFreeHost(h->thetas_host);
  h->results = nullptr; h->mu_in = nullptr; h->results_host = nullptr; h->result_slots = 0;
  h->job_slot_dev = nullptr; h->job_slot_host = nullptr; h->mu_host = nullptr; h->thetas_host = nullptr;
  h->fw.gram_part = nullptr; h->fw.quad_part = nullptr;
  const int pm = h->p + 1;
  VC_CUDA_TRY(cudaMalloc(&h->results, (size_t)m * VC_RES_STRIDE * sizeof(double)));
  VC_CUDA_TRY(cudaMalloc(&h->mu_in, (size_t)m * 128 * sizeof(double)));
  VC_CUDA_TRY(cudaMalloc(&h->job_slot_dev, (size_t)m * sizeof(int)));
  VC_CUDA_TRY(cudaMalloc(&h->fw.gram_part, (size_t)m * h->fw.blocks * pm * (pm + 1) * sizeof(double)));
  VC_CUDA_TRY(cudaMalloc(&h->fw.quad_part, (size_t)m * h->fw.blocks * 2 * sizeof(double)));
  VC_CUDA_TRY(cudaMallocHost(&h->results_host, (size_t)m * VC_RES_STRIDE * sizeof(double)));
  VC_CUDA_TRY(cudaMallocHost(&h->job_slot_host, (size_t)m * sizeof(int)));
  VC_CUDA_TRY(cudaMallocHost(&h->mu_host, (size_t)m * 128 * sizeof(double)));
  // one staging entry per distinct theta of a call (never reused within a call: the H2D copies are asynchronous)
  VC_CUDA_TRY(cudaMallocHost(&h->thetas_host, (size_t)m * sizeof(VcTheta)));
  h->result_slots = m;
}

// per-SLOT buffers: matrix, border, diagonal-tile inverses, log-det partials, info, theta staging
static void alloc_slot_buffers(vc_handle* h, int S) {
  vc_chol_free_graphs(h->cw);            // captured graphs hold the old addresses
  cudaFree(h->M.a); cudaFree(h->M.b); cudaFree(h->cw.linv); cudaFree(h->cw.logdet_part); cudaFree(h->cw.info);
  cudaFree(h->thetas_dev);
  h->M.a = h->M.b = h->cw.linv = h->cw.logdet_part = nullptr;
  h->cw.info = nullptr; h->thetas_dev = nullptr;
  h->slots = 0;
  h->factor_valid = false;
  h->factor_theta.clear();
  const int64_t np = h->n_pad;
  const size_t linv_doubles = (size_t)h->M.nt * VC_TILE * VC_TILE;
  h->M.a_es = h->M.a_doubles;
  h->M.b_es = h->M.ldb * np;
  h->M.nbatch = 1;
  VC_CUDA_TRY(cudaMalloc(&h->M.a, (size_t)S * h->M.a_doubles * sizeof(double)));
  VC_CUDA_TRY(cudaMalloc(&h->M.b, (size_t)S * h->M.b_es * sizeof(double)));
  VC_CUDA_TRY(cudaMalloc(&h->cw.linv, (size_t)S * linv_doubles * sizeof(double)));
  // the diagonal-tile kernel writes only the lower 16x16 blocks of each inverse; the rest stays zero
  // (on the handle's own stream: a legacy-stream memset is not ordered against non-blocking streams)
  VC_CUDA_TRY(cudaMemsetAsync(h->cw.linv, 0, (size_t)S * linv_doubles * sizeof(double), h->st_main));
  VC_CUDA_TRY(cudaMalloc(&h->cw.logdet_part, (size_t)S * h->M.nt * sizeof(double)));
  VC_CUDA_TRY(cudaMalloc(&h->cw.info, (size_t)S * sizeof(int)));
  VC_CUDA_TRY(cudaMalloc(&h->thetas_dev, (size_t)S * sizeof(VcTheta)));
  h->slots = S;
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
    h->cw.nb_tiles = (cfg->nb_outer > 0 ? cfg->nb_outer : auto_outer_block(np)) / VC_TILE;
    h->cw.lookahead = !(cfg->flags & VC_FLAG_NO_LOOKAHEAD);
    h->cw.st_main = h->st_main; h->cw.st_panel = h->st_panel;
    if (alloc_matrix) {
      // lower block-column trapezoids only: block column c starts at row tile c * blk
      const int blk = h->cw.nb_tiles, nt = h->M.nt;
      std::vector<int> r0, ncols;
      for (int c = 0; c * blk < nt; ++c) { r0.push_back(c * blk); ncols.push_back(std::min(blk, nt - c * blk)); }
      vc_set_layout(h, r0, ncols, nt);
      alloc_slot_buffers(h, 1);
    } else {
      h->M.nbatch = 1;
      VC_CUDA_TRY(cudaMalloc(&h->cw.info, sizeof(int)));
    }
    int blocks = (int)(n / 512);
    if (blocks < 1) blocks = 1;
    if (blocks > 148) blocks = 148;
    h->fw.blocks = blocks;
    const int m = p + 1;
    ensure_result_slots(h, 64);
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

// One process, several GPUs: a child handle per device, each with a replica of the O(n) inputs.  vc_n2ll_batch
// spreads the distinct thetas of a batch over the devices -- the GPU analogue of the reference's optimParallel
// fan-out over PSOCK workers (R-pkg/R/SVC_mle.R:164-200); no collective, results are gathered on the host.
extern "C" int vc_create_multi(vc_handle** out, const vc_config* cfg_in, int64_t n, int32_t d, int32_t p,
                               int32_t q, const double* locs, const double* X, const double* W,
                               const double* y, const int32_t* devices, int32_t ndev) {
  if (!out) return VC_ERR_INVALID;
  *out = nullptr;
  if (!devices || ndev < 1 || ndev > 64) return vc_fail(nullptr, VC_ERR_INVALID, "device list must hold 1..64 ordinals");
  vc_config cfg;
  if (cfg_in) cfg = *cfg_in; else vc_config_default(&cfg);
  vc_handle* h = new vc_handle();
  h->cfg = cfg; h->n = n; h->d = d; h->p = p; h->q = q;
  for (int i = 0; i < ndev; ++i) {
    vc_config c = cfg;
    c.device = devices[i];
    vc_handle* k = nullptr;
    const int rc = vc_create(&k, &c, n, d, p, q, locs, X, W, y);
    if (rc != VC_OK) {                        // g_create_error already holds the child's message
      vc_free_handle(h);
      return rc;
    }
    h->kids.push_back(k);
  }
  h->device = devices[0];
  h->n_pad = h->kids[0]->n_pad;
  *out = h;
  return VC_OK;
}

static int multi_batch(vc_handle* h, int32_t m, const double* thetas, int32_t mu_mode, const double* mus_in,
                       double* out_n2ll, double* out_mu, int32_t* out_status) {
  const int nth = 2 * h->q + 1, p = h->p;
  const int G = (int)h->kids.size();
  h->last_info = 0;
  h->err.clear();
  // distinct thetas, in order of first appearance; all evaluations of one theta go to the same device
  std::vector<int> uniq, uniq_of(m);
  for (int e = 0; e < m; ++e) {
    int u = -1;
    for (size_t k = 0; k < uniq.size() && u < 0; ++k)
      if (memcmp(thetas + (size_t)uniq[k] * nth, thetas + (size_t)e * nth, nth * sizeof(double)) == 0) u = (int)k;
    if (u < 0) { u = (int)uniq.size(); uniq.push_back(e); }
    uniq_of[e] = u;
  }
  const int nu = (int)uniq.size();
  // contiguous blocks of distinct thetas per device (neighbouring stencil points have similar cost)
  std::vector<std::vector<int>> evals(G);
  for (int e = 0; e < m; ++e) evals[(int)((int64_t)uniq_of[e] * G / nu)].push_back(e);
  std::vector<int> rcs(G, VC_OK);
  std::vector<std::vector<double>> th(G), mu(G), val(G), muo(G);
  std::vector<std::vector<int32_t>> st(G);
  std::vector<std::thread> workers;
  bool spawn_failed = false;
  for (int g = 0; g < G; ++g) {
    const int mg = (int)evals[g].size();
    if (mg == 0) continue;
    th[g].resize((size_t)mg * nth); val[g].resize(mg); st[g].resize(mg); muo[g].resize((size_t)mg * p);
    if (mus_in) mu[g].resize((size_t)mg * p);
    for (int i = 0; i < mg; ++i) {
      memcpy(&th[g][(size_t)i * nth], thetas + (size_t)evals[g][i] * nth, nth * sizeof(double));
      if (mus_in) memcpy(&mu[g][(size_t)i * p], mus_in + (size_t)evals[g][i] * p, p * sizeof(double));
    }
    auto work = [&, g, mg]() {
      rcs[g] = vc_n2ll_batch(h->kids[g], mg, th[g].data(), mu_mode, mus_in ? mu[g].data() : nullptr, val[g].data(),
                             muo[g].data(), st[g].data());
    };
    try {
      workers.emplace_back(work);            // one host thread per device: vc_n2ll_batch blocks until its results are back
    } catch (const std::exception&) {        // no thread to be had: run this device's share on the calling thread
      spawn_failed = true;
      work();
    }
  }
  for (auto& w : workers) w.join();
  (void)spawn_failed;
  int first = VC_OK;
  for (int g = 0; g < G; ++g) {
    const int mg = (int)evals[g].size();
    if (rcs[g] == VC_ERR_CUDA || rcs[g] == VC_ERR_OOM) return vc_fail(h, rcs[g], h->kids[g]->err);
    for (int i = 0; i < mg; ++i) {
      const int e = evals[g][i];
      out_n2ll[e] = val[g][i];
      if (out_mu) memcpy(out_mu + (size_t)e * p, &muo[g][(size_t)i * p], p * sizeof(double));
      if (out_status) out_status[e] = st[g][i];
    }
    if (mg > 0 && evals[g].back() == m - 1) h->kid_last = g;
  }
  for (int e = 0; e < m && first == VC_OK; ++e) {          // first non-OK status in evaluation order
    const int g = (int)((int64_t)uniq_of[e] * G / nu);
    for (size_t i = 0; i < evals[g].size(); ++i)
      if (evals[g][i] == e && st[g][i] != VC_OK) { first = st[g][i]; h->err = h->kids[g]->err; h->last_info = h->kids[g]->last_info; }
  }
  return first;
}

extern "C" int vc_set_data(vc_handle* h, const double* locs, const double* X, const double* W,
                           const double* y) {
  if (!h) return VC_ERR_INVALID;
  if (!h->kids.empty()) {
    for (vc_handle* k : h->kids) {
      const int rc = vc_set_data(k, locs, X, W, y);
      if (rc != VC_OK) return vc_fail(h, rc, k->err);
    }
    return VC_OK;
  }
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    if (locs) upload_padded(h->locs, locs, h->n, h->n_pad, h->d, h->st_main);
    if (X) upload_padded(h->x, X, h->n, h->n_pad, h->p, h->st_main);
    if (W) upload_padded(h->w, W, h->n, h->n_pad, h->q, h->st_main);
    if (y) upload_padded(h->y, y, h->n, h->n_pad, 1, h->st_main);
    h->factor_valid = false;
    h->factor_theta.clear();
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

// Test hook (host only): the lower block-column storage table of a single-GPU matrix of order n with outer block
// nb_outer (0 = the automatic choice): per block column its offset (doubles), leading dimension and first stored
// row tile.  Returns the number of block columns; *doubles_out = size of one evaluation slot.
extern "C" int vc_layout_describe(int64_t n, int32_t nb_outer, int64_t* doubles_out, int64_t* off, int32_t* ld,
                                  int32_t* r0, int32_t cap) {
  if (n < 1 || nb_outer < 0 || nb_outer % VC_TILE != 0) return -1;
  const int64_t np = (n + VC_TILE - 1) / VC_TILE * VC_TILE;
  const int nt = (int)(np / VC_TILE);
  const int blk = (nb_outer > 0 ? nb_outer : auto_outer_block(np)) / VC_TILE;
  std::vector<int> r0v, ncols;
  for (int c = 0; c * blk < nt; ++c) { r0v.push_back(c * blk); ncols.push_back(std::min(blk, nt - c * blk)); }
  std::vector<VcColBlock> cb;
  const int64_t total = vc_layout_fill(cb, r0v, ncols, nt);
  if (doubles_out) *doubles_out = total;
  for (int c = 0; c < (int)cb.size() && c < cap; ++c) {
    if (off) off[c] = cb[c].off;
    if (ld) ld[c] = cb[c].ld;
    if (r0) r0[c] = cb[c].r0;
  }
  return (int)cb.size();
}

// ---- host-side plan of a batch (no CUDA: unit-tested on the CPU through vc_plan_batch) ----------------------
// Distinct thetas (optim's probes of the mean parameters in the non-profile objective repeat theta: they share
// one factorisation and differ only in the O(n p) finalize job), processed in groups of at most S slots; the theta
// of the last valid evaluation goes LAST so that it lands in slot 0 of the last group (vc_get_chol /
// vc_get_whitened / vc_post refer to slot 0).
struct VcBatchPlan {
  std::vector<int> uniq;      // first evaluation of every distinct theta
  std::vector<int> uniq_of;   // evaluation -> distinct index (-1: invalid theta)
  std::vector<int> order;     // distinct thetas in processing order
  int last_valid = -1;
};
static void plan_batch(int m, int nth, const double* thetas, const std::vector<int>& status, VcBatchPlan& pl) {
  pl.uniq.clear(); pl.order.clear();
  pl.uniq_of.assign(m, -1);
  pl.last_valid = -1;
  for (int e = 0; e < m; ++e) {
    if (status[e] != VC_OK) continue;
    int u = -1;
    for (size_t k = 0; k < pl.uniq.size() && u < 0; ++k)
      if (memcmp(thetas + (size_t)pl.uniq[k] * nth, thetas + (size_t)e * nth, nth * sizeof(double)) == 0) u = (int)k;
    if (u < 0) { u = (int)pl.uniq.size(); pl.uniq.push_back(e); }
    pl.uniq_of[e] = u;
    pl.last_valid = e;
  }
  const int nu = (int)pl.uniq.size();
  const int u_last = pl.last_valid >= 0 ? pl.uniq_of[pl.last_valid] : -1;
  for (int u = 0; u < nu; ++u) if (u != u_last) pl.order.push_back(u);
  if (u_last >= 0) pl.order.push_back(u_last);
}
// slot i of the group order[g0 .. g0 + gsz) holds distinct theta order[g0 + gsz - 1 - i]: its last theta sits in slot 0
static void group_slots(const VcBatchPlan& pl, int g0, int gsz, std::vector<int>& slot_of_u) {
  slot_of_u.assign(pl.uniq.size(), -1);
  for (int i = 0; i < gsz; ++i) slot_of_u[pl.order[g0 + gsz - 1 - i]] = i;
}
// Test hook (host only): for every evaluation its group, slot and result record under S slots; -1 for invalid ones.
extern "C" int vc_plan_batch(int32_t m, int32_t nth, const double* thetas, const int32_t* valid, int32_t S,
                             int32_t* group_of, int32_t* slot_of, int32_t* rec_of) {
  if (m < 1 || nth < 1 || !thetas || S < 1 || !group_of || !slot_of || !rec_of) return -1;
  std::vector<int> status(m, VC_OK);
  for (int e = 0; e < m; ++e) if (valid && !valid[e]) status[e] = VC_ERR_INVALID;
  VcBatchPlan pl;
  plan_batch(m, nth, thetas, status, pl);
  const int nu = (int)pl.uniq.size();
  for (int e = 0; e < m; ++e) group_of[e] = slot_of[e] = rec_of[e] = -1;
  int rec_base = 0, g = 0;
  for (int g0 = 0; g0 < nu; g0 += S, ++g) {
    const int gsz = std::min((int)S, nu - g0);
    std::vector<int> slot_of_u;
    group_slots(pl, g0, gsz, slot_of_u);
    int njobs = 0;
    for (int e = 0; e < m; ++e) {
      if (status[e] != VC_OK || slot_of_u[pl.uniq_of[e]] < 0) continue;
      group_of[e] = g; slot_of[e] = slot_of_u[pl.uniq_of[e]]; rec_of[e] = rec_base + njobs;
      ++njobs;
    }
    rec_base += njobs;
  }
  return nu;
}

// how many evaluation slots a batch may use: one evaluation fills the GPU from n ~ 25k on
static int slot_cap(const vc_handle* h) {
  if (h->dist) return 1;
  static const int env_cap = getenv("VC_MAX_SLOTS") ? atoi(getenv("VC_MAX_SLOTS")) : 0;   // tuning / debugging
  if (h->n_pad > 24576) return 1;
  if (h->slot_cap_cached > 0) return h->slot_cap_cached;
  const double bytes = 8.0 * ((double)h->M.a_doubles + (double)h->M.ldb * h->n_pad + (double)h->M.nt * VC_TILE * VC_TILE);
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const double budget = std::min(24e9, 0.5 * ((double)free_b + (double)h->slots * bytes));
  int cap = (int)std::max(1.0, std::min(32.0, floor(budget / bytes)));
  if (env_cap >= 1) cap = std::min(cap, env_cap);
  const_cast<vc_handle*>(h)->slot_cap_cached = cap;
  return cap;
}

// grow the slot buffers to S evaluations (halving on out-of-memory); returns the slots available
static int ensure_slots(vc_handle* h, int S) {
  if (S <= h->slots) return h->slots;
  while (S > h->slots) {
    try {
      alloc_slot_buffers(h, S);
      return h->slots;
    } catch (const VcError& e) {
      if (e.code != VC_ERR_OOM) throw;
      (void)cudaGetLastError();
      S = std::max(S / 2, 1);
      if (S <= 1) { alloc_slot_buffers(h, 1); return 1; }
    }
  }
  return h->slots;
}

extern "C" int vc_n2ll_batch(vc_handle* h, int32_t m, const double* thetas, int32_t mu_mode,
                             const double* mus_in, double* out_n2ll, double* out_mu,
                             int32_t* out_status) {
  if (!h) return VC_ERR_INVALID;
  if (m < 1 || !thetas || !out_n2ll) return vc_fail(h, VC_ERR_INVALID, "bad batch arguments");
  if (mu_mode != VC_MU_GLS && mu_mode != VC_MU_FIXED) return vc_fail(h, VC_ERR_INVALID, "unknown mu_mode");
  if (mu_mode == VC_MU_FIXED && !mus_in) return vc_fail(h, VC_ERR_INVALID, "mu_mode FIXED needs mu");
  if (!h->kids.empty()) return multi_batch(h, m, thetas, mu_mode, mus_in, out_n2ll, out_mu, out_status);
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
    for (int e = 0; e < m; ++e) {
      VcTheta th;
      if (vc_make_theta(h, thetas + (size_t)e * nth, &th) != VC_OK) status[e] = VC_ERR_INVALID;
    }
    VcBatchPlan bp;
    plan_batch(m, nth, thetas, status, bp);
    const std::vector<int>& uniq_of = bp.uniq_of;
    const std::vector<int>& uniq = bp.uniq;
    const int last_valid = bp.last_valid;
    const int nu = (int)uniq.size();
    // the resident factor of slot 0 (a repeated theta, vc_post after a fit) is reused when the whole call is about it
    const bool cached = nu == 1 && h->factor_theta.size() == (size_t)nth &&
                        memcmp(h->factor_theta.data(), thetas + (size_t)uniq[0] * nth, nth * sizeof(double)) == 0;
    int S = 1;
    if (nu > 1) S = ensure_slots(h, std::min(nu, slot_cap(h)));
    // order of the distinct thetas: the theta of the last valid evaluation goes last, so that it lands in slot 0
    // of the last group (vc_get_chol / vc_get_whitened / vc_post refer to slot 0)
    const std::vector<int>& order = bp.order;
    std::vector<int> rec_of(m, -1);             // evaluation -> result record
    int rec_base = 0;
    cudaStream_t st = h->st_main;
    for (int g0 = 0; g0 < nu; g0 += S) {
      const int gsz = std::min(S, nu - g0);
      const bool last_group = g0 + gsz >= nu;
      // slot i of this group holds distinct theta order[g0 + gsz - 1 - i]: the group's last theta sits in slot 0
      std::vector<int> slot_of_u;
      group_slots(bp, g0, gsz, slot_of_u);
      for (int i = 0; i < gsz; ++i)
        vc_make_theta(h, thetas + (size_t)uniq[order[g0 + gsz - 1 - i]] * nth, &h->thetas_host[g0 + i]);
      // jobs of this group, in evaluation order
      int njobs = 0;
      for (int e = 0; e < m; ++e) {
        if (status[e] != VC_OK || slot_of_u[uniq_of[e]] < 0) continue;
        const int r = rec_base + njobs;
        rec_of[e] = r;
        h->job_slot_host[r] = slot_of_u[uniq_of[e]];
        if (mu_mode == VC_MU_FIXED)
          memcpy(h->mu_host + (size_t)r * 128, mus_in + (size_t)e * h->p, h->p * sizeof(double));
        ++njobs;
      }
      if (njobs == 0) continue;
      VC_CUDA_TRY(cudaMemcpyAsync(h->job_slot_dev + rec_base, h->job_slot_host + rec_base, njobs * sizeof(int),
                                  cudaMemcpyHostToDevice, st));
      if (mu_mode == VC_MU_FIXED)
        VC_CUDA_TRY(cudaMemcpyAsync(h->mu_in + (size_t)rec_base * 128, h->mu_host + (size_t)rec_base * 128,
                                    (size_t)njobs * 128 * sizeof(double), cudaMemcpyHostToDevice, st));
      VcMatrix M = h->M;
      M.nbatch = gsz;
      M.nbt = 1;
      const bool timed = last_group;
      if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[0], st));
      if (!cached) {
        if (gsz > 1) {
          VC_CUDA_TRY(cudaMemcpyAsync(h->thetas_dev, h->thetas_host + g0, gsz * sizeof(VcTheta), cudaMemcpyHostToDevice, st));
          vc_launch_build(M, h->thetas_host[g0], h->d, h->locs, h->w, st, &h->launches, h->thetas_dev);
        } else {
          vc_launch_build(M, h->thetas_host[g0], h->d, h->locs, h->w, st, &h->launches);   // theta by value: no copy
        }
        vc_launch_border(M, h->p, h->x, h->y, st, &h->launches);
      }
      if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[1], st));
      if (!cached) {
        h->factor_theta.clear();
        vc_chol_factor(M, h->cw, &h->launches);
      }
      if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[2], st));
      vc_launch_finalize(M, h->p, mu_mode, h->mu_in + (size_t)rec_base * 128, h->cw.logdet_part, h->cw.info, h->fw,
                         h->results + (size_t)rec_base * VC_RES_STRIDE, st, &h->launches, h->gram_dev,
                         njobs, h->job_slot_dev + rec_base);
      if (timed) VC_CUDA_TRY(cudaEventRecord(h->ev[3], st));
      rec_base += njobs;
    }
    if (rec_base > 0)
      VC_CUDA_TRY(cudaMemcpyAsync(h->results_host, h->results, (size_t)rec_base * VC_RES_STRIDE * sizeof(double),
                                  cudaMemcpyDeviceToHost, st));
    VC_CUDA_TRY(cudaStreamSynchronize(st));
    VC_CUDA_TRY(cudaGetLastError());
    for (int e = 0; e < m; ++e) {
      if (status[e] == VC_OK) {
        const double* r = h->results_host + (size_t)rec_of[e] * VC_RES_STRIDE;
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
    if (last_valid >= 0) {
      float b = 0, c = 0, f = 0;
      cudaEventElapsedTime(&b, h->ev[0], h->ev[1]);
      cudaEventElapsedTime(&c, h->ev[1], h->ev[2]);
      cudaEventElapsedTime(&f, h->ev[2], h->ev[3]);
      h->last_timing = {b, c, f, b + c + f};
      // slot 0 now holds the factor of the last valid evaluation
      h->factor_theta.assign(thetas + (size_t)last_valid * nth, thetas + (size_t)(last_valid + 1) * nth);
      h->factor_valid = (status[last_valid] == VC_OK);
      h->last_rec = rec_of[last_valid];
    }
    if (first == VC_ERR_NOT_PD) {
      char buf[128];
      snprintf(buf, sizeof(buf), "the leading minor of order %d is not positive", h->last_info);
      h->err = buf;
    } else if (first == VC_ERR_INVALID) {
      h->err = "theta does not define a covariance (range <= 0 or non-finite value)";
    }
  } catch (const VcError& e) {
    h->factor_theta.clear();
    h->factor_valid = false;
    return vc_fail(h, e.code, e.msg);
  }
  return first;
}

extern "C" int vc_n2ll(vc_handle* h, const double* theta, int32_t mu_mode, const double* mu_in,
                       double* n2ll, double* mu_out, double* logdet, double* quad) {
  if (!h || !theta || !n2ll) return VC_ERR_INVALID;
  if (!h->kids.empty()) {            // a lone evaluation runs on the child that holds the last factor (cache hits)
    vc_handle* k = vc_target(h);
    const int rc = vc_n2ll(k, theta, mu_mode, mu_in, n2ll, mu_out, logdet, quad);
    h->last_info = k->last_info;
    h->err = rc != VC_OK ? k->err : std::string();
    return rc;
  }
  if (h->dist) return vc_dist_n2ll(h, theta, mu_mode, mu_in, n2ll, mu_out, logdet, quad);
  int32_t status = VC_OK;
  h->last_info = 0;
  int rc = vc_n2ll_batch(h, 1, theta, mu_mode, mu_in, n2ll, mu_out, &status);
  if (rc == VC_OK || rc == VC_ERR_NOT_PD) {
    const double* r = h->results_host + (size_t)h->last_rec * VC_RES_STRIDE;
    if (logdet) *logdet = r[VC_RES_LOGDET];
    if (quad) *quad = r[VC_RES_QUAD];
  }
  return rc;
}

extern "C" int vc_sigma(vc_handle* h, const double* theta, double* out) {
  if (!h || !theta || !out) return VC_ERR_INVALID;
  h = vc_target(h);
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
  h = vc_target(h);
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
  h = vc_target(h);
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
  h = vc_target(h);
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
  h = vc_target(h);
  *out = h->last_timing;
  return VC_OK;
}
extern "C" int vc_last_info(vc_handle* h) {
  if (!h) return 0;
  return (!h->kids.empty() && h->last_info == 0) ? vc_target(h)->last_info : h->last_info;
}
extern "C" const char* vc_last_error(vc_handle* h) {
  if (!h) return g_create_error.c_str();
  return (!h->kids.empty() && h->err.empty()) ? vc_target(h)->err.c_str() : h->err.c_str();
}
extern "C" int64_t vc_padded_n(vc_handle* h) { return h ? h->n_pad : 0; }
extern "C" int64_t vc_launch_count(vc_handle* h) {
  if (!h) return 0;
  int64_t s = h->launches;
  for (vc_handle* k : h->kids) s += k->launches;
  return s;
}
extern "C" int32_t vc_device_count(vc_handle* h) { return h ? (h->kids.empty() ? 1 : (int32_t)h->kids.size()) : 0; }
