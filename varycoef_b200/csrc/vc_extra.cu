// vc_extra.cu -- the rows next to the hot path, all on the resident factor:
//   vc_post      prep_par_output + eff_dof   (reference R-pkg/R/utils.R:423-473, R-pkg/R/eff_dof.R:7-21)
//   vc_predict   predict.SVC_mle             (reference R-pkg/R/predict-SVC_mle.R:80-267)
//   vc_gls_chol  GLS_chol.matrix             (reference R-pkg/R/utils.R:93-101, exported)
// Every triangular solve is the blocked forward substitution of vc_chol.cu (VC_SWEEP_SOLVE): the
// right-hand sides are loaded as border rows and swept with the stored diagonal-tile inverses.
#include "vc_handle.cuh"
#include "vc_math.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {

constexpr int TILE = VC_TILE;
__constant__ double c_exp2_tab_x[VC_EXP_TABLE_SIZE];

void extra_init() {
  static bool done_dev[64] = {false};
  int dev = 0;
  VC_CUDA_TRY(cudaGetDevice(&dev));
  if (done_dev[dev & 63]) return;
  double tab[VC_EXP_TABLE_SIZE];
  vc_fill_exp2_table(tab);
  VC_CUDA_TRY(cudaMemcpyToSymbol(c_exp2_tab_x, tab, sizeof(tab)));
  done_dev[dev & 63] = true;
}

// ---- identity border, L -> matrix upload ------------------------------------------------------
__global__ void k_set_identity(double* b, int64_t ldb, int64_t n_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) b[i + i * ldb] = 1.0;
}

// a(i, j) = R(j, i) for i >= j (L = R^T), unit diagonal / zeros in the padding
__global__ void k_load_factor(const double* __restrict__ R, int64_t n, double* __restrict__ a,
                              const VcColBlock* __restrict__ cb, int blk, int64_t n_pad) {
  __shared__ double t[32][33];
  const int64_t bi = (int64_t)blockIdx.x * 32, bj = (int64_t)blockIdx.y * 32;   // block of a: rows bi.., cols bj..
  if (bi + 31 < bj) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int yy = ty; yy < 32; yy += 8) {          // read R(bj + yy, bi + tx): row index fast? R col-major: R[r + c n]
    const int64_t r = bj + tx, c = bi + yy;      // R(r, c), r fast
    t[yy][tx] = (r < n && c < n) ? R[r + c * n] : 0.0;
  }
  __syncthreads();
  for (int yy = ty; yy < 32; yy += 8) {
    const int64_t i = bi + tx, j = bj + yy;      // a(i, j) = R(j, i) = t[i - bi][j - bj]
    if (i < n_pad && j < n_pad && i >= j) {
      double v = t[tx][yy];
      if (i >= n || j >= n) v = (i == j) ? 1.0 : 0.0;
      *vc_elem_ptr(a, cb, blk, i, j) = v;
    }
  }
}

// ---- eff_dof pieces: per row i of L^-T (stored in the identity border): f_i = |row|^2, V_i = row * Z_X ----
template <int PMAX>
__global__ void __launch_bounds__(128) k_edof_rows(const double* __restrict__ b, int64_t ldb, int64_t n, int p,
                                                   const double* __restrict__ zx, double* __restrict__ f,
                                                   double* __restrict__ V) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[PMAX];
#pragma unroll
  for (int c = 0; c < PMAX; ++c) v[c] = 0.0;
  double fs = 0.0;
  for (int64_t j = i; j < n; ++j) {
    const double x = b[i + j * ldb];
    fs = fma(x, x, fs);
#pragma unroll
    for (int c = 0; c < PMAX; ++c)
      if (c < p) v[c] = fma(x, zx[j + (int64_t)c * n], v[c]);
  }
  f[i] = fs;
#pragma unroll
  for (int c = 0; c < PMAX; ++c)
    if (c < p) V[i + (int64_t)c * n] = v[c];
}

// ---- predict: residual in whitened space, back substitution, cross-covariance products ----------
// rz[j] = Z_y[j] - sum_c Z_X[j, c] mu[c]   (border tile 0, rows 0..p)
__global__ void k_resid(const double* __restrict__ b, int64_t ldb, int64_t n_pad, int p, const double* __restrict__ mu,
                        double* __restrict__ rz) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pad) return;
  const double* z = b + j * ldb;
  double r = z[p];
  for (int c = 0; c < p; ++c) r = fma(-z[c], mu[c], r);
  rz[j] = r;
}
// alpha_jt = Linv_jt^T r_jt  (in place on r)
__global__ void __launch_bounds__(128) k_bs_diag(const double* __restrict__ linv, double* __restrict__ r) {
  __shared__ double rs[TILE];
  const int c = threadIdx.x;
  rs[c] = r[c];
  __syncthreads();
  double s = 0.0;
  for (int i = c; i < TILE; ++i) s = fma(linv[i + TILE * c], rs[i], s);
  r[c] = s;
}
// r_k -= L(jt, k)^T alpha_jt for every k < jt (one CTA per k)
__global__ void __launch_bounds__(128) k_bs_update(double* a, const VcColBlock* __restrict__ cb, int blk, int jt,
                                                   double* __restrict__ r) {
  __shared__ double al[TILE];
  const int k = blockIdx.x, c = threadIdx.x;
  al[c] = r[(int64_t)jt * TILE + c];
  __syncthreads();
  int64_t lda;
  const double* col = vc_tile_ptr(a, cb, blk, jt, k, lda) + (int64_t)c * lda;
  double s = 0.0;
#pragma unroll 8
  for (int i = 0; i < TILE; ++i) s = fma(col[i], al[i], s);
  r[(int64_t)k * TILE + c] -= s;
}

struct CrossArgs {
  const double* newlocs;   // n_new x d column-major
  int64_t n_new;
  const double* locs;      // n_pad x d
  const double* w;         // n_pad x q
  int64_t n, n_pad;
  int d;
  VcTheta th;
};
__device__ __forceinline__ double pair_dist(const CrossArgs& p, const double* xi, const double* sx, int jj) {
  double s = 0.0;
  for (int c = 0; c < p.d; ++c) s = vc_dist_accum(p.th.dist_method, s, xi[c] - sx[c * TILE + jj], p.th.mink_p);
  return vc_dist_finish(p.th.dist_method, s, p.th.mink_p);
}

// eff partial: part[(chunk * n_new + i) * q + k] = sum_{j in chunk} cov_k(h_ij) taper(h_ij) W_jk alpha_j
template <int QMAX>
__global__ void __launch_bounds__(128) k_cross_eff(CrossArgs p, const double* __restrict__ alpha, int chunk_len,
                                                   double* __restrict__ part) {
  extern __shared__ double sm[];
  double* tab = sm;
  double* sx = tab + VC_EXP_TABLE_SIZE;          // d x 128
  double* sw = sx + p.d * TILE;                  // q x 128 (W_jk * alpha_j)
  const int q = p.th.q;
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * TILE + tid;
  if (tid < VC_EXP_TABLE_SIZE) tab[tid] = c_exp2_tab_x[tid];
  double xi[VC_MAXD];
  for (int c = 0; c < p.d; ++c) xi[c] = (i < p.n_new) ? p.newlocs[i + (int64_t)c * p.n_new] : 0.0;
  double acc[QMAX];
#pragma unroll
  for (int k = 0; k < QMAX; ++k) acc[k] = 0.0;
  const int64_t j0 = (int64_t)blockIdx.y * chunk_len;
  const int64_t j1 = min(j0 + (int64_t)chunk_len, p.n);
  for (int64_t jb = j0; jb < j1; jb += TILE) {
    __syncthreads();
    const int64_t j = jb + tid;
    const bool ok = j < j1;
    for (int c = 0; c < p.d; ++c) sx[c * TILE + tid] = ok ? p.locs[j + (int64_t)c * p.n_pad] : 0.0;
    const double aj = ok ? alpha[j] : 0.0;
    for (int k = 0; k < q; ++k) sw[k * TILE + tid] = ok ? p.w[j + (int64_t)k * p.n_pad] * aj : 0.0;
    __syncthreads();
    const int lim = (int)min((int64_t)TILE, j1 - jb);
    for (int jj = 0; jj < lim; ++jj) {
      const double h = pair_dist(p, xi, sx, jj);
      const double tp = vc_taper(p.th.taper_id, h, p.th.inv_delta);
#pragma unroll
      for (int k = 0; k < QMAX; ++k)
        if (k < q)
          acc[k] = fma(p.th.sill[k] * vc_cov_shape_rt(p.th.cov_id, h * p.th.inv_range[k], tab) * tp,
                       sw[k * TILE + jj], acc[k]);
    }
  }
  if (i < p.n_new)
#pragma unroll
    for (int k = 0; k < QMAX; ++k)
      if (k < q) part[((int64_t)blockIdx.y * p.n_new + i) * q + k] = acc[k];
}
__global__ void k_reduce_chunks(const double* __restrict__ part, int chunks, int64_t count, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += part[(int64_t)c * count + i];
  out[i] = s;
}

// border rows = cross-covariance of y_new with y (Sigma_y_y, reference R-pkg/R/utils.R:252-269):
// B(i, j) = sum_k cov_k(h_ij) taper(h_ij) newW_ik W_jk   for the rows [i0, i0 + 128 nbt) of newlocs
__global__ void __launch_bounds__(256) k_cross_border(CrossArgs p, const double* __restrict__ neww, int64_t i0,
                                                      double* __restrict__ b, int64_t ldb) {
  extern __shared__ double sm[];
  double* tab = sm;
  double* sx = tab + VC_EXP_TABLE_SIZE;          // d x 128 (train locs of this column tile)
  double* sw = sx + p.d * TILE;                  // q x 128
  const int q = p.th.q;
  const int tid = threadIdx.x;
  if (tid < VC_EXP_TABLE_SIZE) tab[tid] = c_exp2_tab_x[tid];
  const int64_t jb = (int64_t)blockIdx.y * TILE;
  for (int idx = tid; idx < p.d * TILE; idx += blockDim.x)
    sx[idx] = p.locs[(jb + idx % TILE) + (int64_t)(idx / TILE) * p.n_pad];
  for (int idx = tid; idx < q * TILE; idx += blockDim.x)
    sw[idx] = p.w[(jb + idx % TILE) + (int64_t)(idx / TILE) * p.n_pad];
  __syncthreads();
  const int r = tid & 127, half = tid >> 7;
  const int64_t lr = (int64_t)blockIdx.x * TILE + r;   // row inside the border
  const int64_t i = i0 + lr;
  double xi[VC_MAXD];
  for (int c = 0; c < p.d; ++c) xi[c] = (i < p.n_new) ? p.newlocs[i + (int64_t)c * p.n_new] : 0.0;
  for (int jj = half * 64; jj < half * 64 + 64; ++jj) {
    double v = 0.0;
    if (i < p.n_new && jb + jj < p.n) {
      const double h = pair_dist(p, xi, sx, jj);
      const double tp = vc_taper(p.th.taper_id, h, p.th.inv_delta);
      for (int k = 0; k < q; ++k)
        v = fma(p.th.sill[k] * vc_cov_shape_rt(p.th.cov_id, h * p.th.inv_range[k], tab) * tp,
                neww[i + (int64_t)k * p.n_new] * sw[k * TILE + jj], v);
    }
    b[lr + (jb + jj) * ldb] = v;
  }
}
__global__ void __launch_bounds__(128) k_row_sqnorm(const double* __restrict__ b, int64_t ldb, int64_t rows,
                                                    int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double s = 0.0;
  for (int64_t j = 0; j < n; ++j) {
    const double x = b[i + j * ldb];
    s = fma(x, x, s);
  }
  out[i] = s;
}

// p x p symmetric positive definite inverse (Cholesky), column-major, host
bool spd_inverse(int p, const double* g, int ldg, double* out) {
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

struct DevBuf {   // RAII device allocation
  double* p = nullptr;
  explicit DevBuf(size_t n) { VC_CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double))); }
  ~DevBuf() { cudaFree(p); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// run a solve sweep over `nbt` border tiles stored in xb (ld = 128 nbt), leaving the handle's own border intact
void sweep_border(vc_handle* h, double* xb, int nbt, int mode) {
  VcMatrix M = h->M;
  M.b = xb; M.ldb = (int64_t)TILE * nbt; M.nbt = nbt;
  vc_chol_factor(M, h->cw, &h->launches, mode);
}

// eff_dof (R-pkg/R/eff_dof.R:7-21) from L^-T:  tau^2 tr(Sigma_FE X' S^-2 X) + n - tau^2 tr(S^-1)
int post_edof(vc_handle* h, const double* theta, const std::vector<double>& sigma_fe, double* edof) {
  const int p = h->p;
  const int64_t n = h->n, np = h->n_pad;
  if (p > 32) return vc_fail(h, VC_ERR_INVALID, "edof is implemented for p <= 32");
  cudaStream_t st = h->st_main;
  DevBuf zx((size_t)n * (p + 1)), xb((size_t)np * np), f((size_t)n), V((size_t)n * p);
  vc_launch_gather_whitened(h->M, p, zx.p, st, &h->launches);
  VC_CUDA_TRY(cudaMemsetAsync(xb.p, 0, (size_t)np * np * sizeof(double), st));
  k_set_identity<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(xb.p, np, np);
  sweep_border(h, xb.p, h->M.nt, VC_SWEEP_SOLVE_IDENT);
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (p <= 8) k_edof_rows<8><<<grid, 128, 0, st>>>(xb.p, np, n, p, zx.p, f.p, V.p);
  else k_edof_rows<32><<<grid, 128, 0, st>>>(xb.p, np, n, p, zx.p, f.p, V.p);
  h->launches += 2;
  std::vector<double> fh(n), Vh((size_t)n * p);
  VC_CUDA_TRY(cudaMemcpyAsync(fh.data(), f.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  VC_CUDA_TRY(cudaMemcpyAsync(Vh.data(), V.p, (size_t)n * p * sizeof(double), cudaMemcpyDeviceToHost, st));
  VC_CUDA_TRY(cudaStreamSynchronize(st));
  long double tr_inv = 0.0L;
  for (int64_t i = 0; i < n; ++i) tr_inv += fh[i];
  std::vector<long double> vtv((size_t)p * p, 0.0L);
  for (int a = 0; a < p; ++a)
    for (int b = a; b < p; ++b) {
      long double s = 0.0L;
      for (int64_t i = 0; i < n; ++i) s += (long double)Vh[i + (size_t)a * n] * Vh[i + (size_t)b * n];
      vtv[a + (size_t)p * b] = vtv[b + (size_t)p * a] = s;
    }
  long double tr2 = 0.0L;      // tr(Sigma_FE * V'V)
  for (int a = 0; a < p; ++a)
    for (int b = 0; b < p; ++b) tr2 += (long double)sigma_fe[a + (size_t)p * b] * vtv[b + (size_t)p * a];
  const double tau2 = theta[2 * h->q];
  *edof = (double)((long double)tau2 * tr2 + (long double)n - (long double)tau2 * tr_inv);
  return VC_OK;
}

}  // namespace

extern "C" int vc_post(vc_handle* h, const double* theta, double* mu_gls, double* sigma_fe, double* edof) {
  if (!h || !theta) return VC_ERR_INVALID;
  h = vc_target(h);
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
    if (!spd_inverse(p, gram.data(), m, sfe.data()))
      return vc_fail(h, VC_ERR_NOT_PD, "X' Sigma^-1 X is not positive definite");
    if (mu_gls) memcpy(mu_gls, mu.data(), p * sizeof(double));
    if (sigma_fe) memcpy(sigma_fe, sfe.data(), sfe.size() * sizeof(double));
    if (edof) return post_edof(h, theta, sfe, edof);
  } catch (const VcError& e) {
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}

extern "C" int vc_predict(vc_handle* h, const double* theta, const double* mu, int64_t n_new,
                          const double* newlocs, const double* newX, const double* newW, double* eff,
                          double* y_pred, double* y_var) {
  if (!h || !theta || !mu || !newlocs || !eff || n_new < 1) return VC_ERR_INVALID;
  h = vc_target(h);
  if (h->dist) return vc_fail(h, VC_ERR_INVALID, "vc_predict is not available on a distributed handle");
  if ((y_pred || y_var) && (!newX || !newW)) return vc_fail(h, VC_ERR_INVALID, "y_pred / y_var need newX and newW");
  const int p = h->p, q = h->q, d = h->d;
  const int64_t n = h->n, np = h->n_pad;
  // factor at theta (or reuse the resident one); the value itself is not needed
  double val = 0.0;
  int rc = vc_n2ll(h, theta, VC_MU_FIXED, mu, &val, nullptr, nullptr, nullptr);
  if (rc != VC_OK) return rc;
  VcTheta th;
  if (vc_make_theta(h, theta, &th) != VC_OK) return vc_fail(h, VC_ERR_INVALID, "invalid theta");
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    extra_init();
    cudaStream_t st = h->st_main;
    DevBuf rz((size_t)np), mud((size_t)p), nl((size_t)n_new * d);
    VC_CUDA_TRY(cudaMemcpyAsync(mud.p, mu, p * sizeof(double), cudaMemcpyHostToDevice, st));
    VC_CUDA_TRY(cudaMemcpyAsync(nl.p, newlocs, (size_t)n_new * d * sizeof(double), cudaMemcpyHostToDevice, st));
    // alpha = Sigma^-1 (y - X mu) = L^-T (Z_y - Z_X mu): block back substitution
    k_resid<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(h->M.b, h->M.ldb, np, p, mud.p, rz.p);
    for (int jt = h->M.nt - 1; jt >= 0; --jt) {
      k_bs_diag<<<1, 128, 0, st>>>(h->cw.linv + (size_t)jt * TILE * TILE, rz.p + (size_t)jt * TILE);
      if (jt > 0) k_bs_update<<<jt, 128, 0, st>>>(h->M.a, h->M.cb, h->M.blk, jt, rz.p);
    }
    h->launches += 1 + 2 * h->M.nt;
    // eff[i, k] = sum_j cov_k(d(new_i, s_j)) taper W_jk alpha_j   (Sigma_b_y %*% alpha, R-pkg/R/utils.R:238-249)
    CrossArgs ca{};
    ca.newlocs = nl.p; ca.n_new = n_new; ca.locs = h->locs; ca.w = h->w; ca.n = n; ca.n_pad = np; ca.d = d; ca.th = th;
    const int chunk_len = 4096;
    const int chunks = (int)((n + chunk_len - 1) / chunk_len);
    DevBuf part((size_t)chunks * n_new * q), effd((size_t)n_new * q);
    const size_t smem = (VC_EXP_TABLE_SIZE + (size_t)(d + q) * TILE) * sizeof(double);
    dim3 grid((unsigned)((n_new + TILE - 1) / TILE), (unsigned)chunks);
    if (q <= 8) k_cross_eff<8><<<grid, 128, smem, st>>>(ca, rz.p, chunk_len, part.p);
    else k_cross_eff<VC_MAXQ><<<grid, 128, smem, st>>>(ca, rz.p, chunk_len, part.p);
    k_reduce_chunks<<<(unsigned)(((int64_t)n_new * q + 255) / 256), 256, 0, st>>>(part.p, chunks, (int64_t)n_new * q, effd.p);
    h->launches += 2;
    std::vector<double> effh((size_t)n_new * q);   // device layout: [i][k]
    VC_CUDA_TRY(cudaMemcpyAsync(effh.data(), effd.p, effh.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    VC_CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n_new; ++i)
      for (int k = 0; k < q; ++k) eff[i + (size_t)k * n_new] = effh[(size_t)i * q + k];
    if (y_pred)
      for (int64_t i = 0; i < n_new; ++i) {
        double s = 0.0;
        for (int k = 0; k < q; ++k) s += newW[i + (size_t)k * n_new] * eff[i + (size_t)k * n_new];
        double xm = 0.0;
        for (int c = 0; c < p; ++c) xm += newX[i + (size_t)c * n_new] * mu[c];
        y_pred[i] = s + xm;       // apply(newW * eff, 1, sum) + newX %*% mu
      }
    if (y_var) {
      // var.y = diag(Sigma_ynew) - diag(B Sigma^-1 B'),  B = Sigma_ynew_y: rows of B swept as border rows
      DevBuf nw((size_t)n_new * q);
      VC_CUDA_TRY(cudaMemcpyAsync(nw.p, newW, (size_t)n_new * q * sizeof(double), cudaMemcpyHostToDevice, st));
      const int max_tiles = 32;                       // 4096 new locations per sweep
      const int nbt_max = (int)std::min<int64_t>(max_tiles, (n_new + TILE - 1) / TILE);
      DevBuf xb((size_t)nbt_max * TILE * np), nrm((size_t)nbt_max * TILE);
      std::vector<double> nh((size_t)nbt_max * TILE);
      for (int64_t i0 = 0; i0 < n_new; i0 += (int64_t)nbt_max * TILE) {
        const int nbt = (int)std::min<int64_t>(nbt_max, (n_new - i0 + TILE - 1) / TILE);
        const int64_t ldb = (int64_t)nbt * TILE;
        dim3 g2((unsigned)nbt, (unsigned)h->M.nt);
        k_cross_border<<<g2, 256, smem, st>>>(ca, nw.p, i0, xb.p, ldb);
        sweep_border(h, xb.p, nbt, VC_SWEEP_SOLVE);
        k_row_sqnorm<<<(unsigned)((ldb + 127) / 128), 128, 0, st>>>(xb.p, ldb, ldb, n, nrm.p);
        h->launches += 2;
        VC_CUDA_TRY(cudaMemcpyAsync(nh.data(), nrm.p, ldb * sizeof(double), cudaMemcpyDeviceToHost, st));
        VC_CUDA_TRY(cudaStreamSynchronize(st));
        for (int64_t r = 0; r < ldb && i0 + r < n_new; ++r) {
          const int64_t i = i0 + r;
          double dg = theta[2 * q];                   // nugget on the diagonal of Sigma_ynew
          for (int k = 0; k < q; ++k) dg += theta[2 * k + 1] * newW[i + (size_t)k * n_new] * newW[i + (size_t)k * n_new];
          y_var[i] = dg - nh[r];
        }
      }
    }
  } catch (const VcError& e) {
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}

extern "C" int vc_gls_chol(int32_t device, int64_t n, int32_t p, const double* R, const double* X,
                           const double* y, double* mu_out) {
  if (!R || !X || !y || !mu_out || n < 1 || p < 1 || p + 1 > VC_BORDER_ROWS) return VC_ERR_INVALID;
  vc_config cfg;
  vc_config_default(&cfg);
  cfg.device = device;
  std::vector<double> zeros((size_t)n, 0.0);
  vc_handle* h = nullptr;
  // a handle provides streams, padded X / y, the matrix and the finalize workspace; locs / W are unused
  int rc = vc_create(&h, &cfg, n, 1, p, 1, zeros.data(), X, X, y);
  if (rc != VC_OK) return rc;
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->st_main;
    {
      DevBuf Rd((size_t)n * n);
      VC_CUDA_TRY(cudaMemcpyAsync(Rd.p, R, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, st));
      dim3 grid((unsigned)((h->n_pad + 31) / 32), (unsigned)((h->n_pad + 31) / 32));
      k_load_factor<<<grid, dim3(32, 8), 0, st>>>(Rd.p, n, h->M.a, h->M.cb, h->M.blk, h->n_pad);
      VC_CUDA_TRY(cudaStreamSynchronize(st));
    }
    vc_chol_invert_diag_tiles(h->M, h->cw, &h->launches);
    vc_launch_border(h->M, p, h->x, h->y, st, &h->launches);
    h->M.nbt = 1;
    vc_chol_factor(h->M, h->cw, &h->launches, VC_SWEEP_SOLVE);
    vc_launch_finalize(h->M, p, VC_MU_GLS, h->mu_in, h->cw.logdet_part, h->cw.info, h->fw, h->results, st,
                       &h->launches, h->gram_dev);
    VC_CUDA_TRY(cudaMemcpyAsync(h->results_host, h->results, VC_RES_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, st));
    VC_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < p; ++i) mu_out[i] = h->results_host[VC_RES_MU + i];
    rc = (h->results_host[VC_RES_INFO] != 0.0) ? VC_ERR_INVALID : VC_OK;
  } catch (const VcError& e) {
    rc = e.code;
  }
  vc_destroy(h);
  return rc;
}
