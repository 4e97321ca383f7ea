// vc_finalize.cu -- K3: GLS normal equations, residual quadratic form, log-determinant and the
// final -2 logL scalar, from the whitened border Z^T = (L^-1 [X y])^T left by the Cholesky.
//
// Replaces GLS_chol.matrix (reference R-pkg/R/utils.R:93-101: solve(crossprod(RiX)) %*%
// crossprod(RiX, Riy)), res <- y - X mu, crossprod(solve(t(cholS), res)) and
// n log(2 pi) + 2 determinant(cholS)$modulus + quad  (R-pkg/R/objective_functions.R:31-61).
// By linearity L^-1 (y - X mu) = Z_y - Z_X mu, so all three mu modes need only Z.
// All reductions run in a fixed order with error-free (two-sum / two-prod) accumulation, so the
// result is deterministic and the O(n) sums carry ~1e-32 relative accumulation error.
#include "vc_common.cuh"
#include <limits.h>
#include <math.h>
#include <algorithm>

namespace {

constexpr int FIN_THREADS = 256;
constexpr int CHUNK = 64;   // columns of Z^T staged per step

// (hi, lo) += a * b, error-free
__device__ __forceinline__ void dd_fma_acc(double& hi, double& lo, double a, double b) {
  const double pr = a * b;
  const double pe = fma(a, b, -pr);
  const double s = hi + pr;
  const double bb = s - hi;
  const double e = (hi - (s - bb)) + (pr - bb);
  hi = s;
  lo += e + pe;
}
__device__ __forceinline__ void dd_add(double& hi, double& lo, double ah, double al) {
  const double s = hi + ah;
  const double bb = s - hi;
  const double e = (hi - (s - bb)) + (ah - bb);
  hi = s;
  lo += e + al;
}

// partial Gram of the (p+1) border rows over a slice of columns; pairs (a <= b) are linearised
__global__ void __launch_bounds__(FIN_THREADS)
k_gram_partial(const double* __restrict__ b, int64_t ldb, int64_t n, int p,
               double* __restrict__ part, int64_t b_es, const int* __restrict__ job_slot) {
  extern __shared__ double sm[];
  const int m = p + 1;
  const int npairs = m * (m + 1) / 2;
  b += (job_slot ? job_slot[blockIdx.y] : 0) * b_es;        // job blockIdx.y reads the border of its slot
  part += (int64_t)blockIdx.y * gridDim.x * 2 * npairs;
  double* z = sm;                         // m x CHUNK
  double* acc = z + m * CHUNK;            // npairs x 2
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * npairs; i += FIN_THREADS) acc[i] = 0.0;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t j0 = (int64_t)blockIdx.x * per;
  int64_t j1 = j0 + per;
  if (j1 > n) j1 = n;
  for (int64_t jb = j0; jb < j1; jb += CHUNK) {
    __syncthreads();
    for (int i = tid; i < m * CHUNK; i += FIN_THREADS) {
      const int c = i % m, jj = i / m;
      const int64_t j = jb + jj;
      z[c * CHUNK + jj] = (j < j1) ? b[c + j * ldb] : 0.0;
    }
    __syncthreads();
    for (int pi = tid; pi < npairs; pi += FIN_THREADS) {
      // decode pair index -> (ra <= rb)
      int rb = (int)((sqrt(8.0 * pi + 1.0) - 1.0) * 0.5);
      while ((rb + 1) * (rb + 2) / 2 <= pi) ++rb;
      while (rb * (rb + 1) / 2 > pi) --rb;
      const int ra = pi - rb * (rb + 1) / 2;
      double hi = acc[2 * pi], lo = acc[2 * pi + 1];
      for (int jj = 0; jj < CHUNK; ++jj) dd_fma_acc(hi, lo, z[ra * CHUNK + jj], z[rb * CHUNK + jj]);
      acc[2 * pi] = hi;
      acc[2 * pi + 1] = lo;
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * npairs; i += FIN_THREADS) part[(int64_t)blockIdx.x * 2 * npairs + i] = acc[i];
}

// one CTA: reduce the partial Grams in block order, solve the p x p normal equations (GLS) or
// take the caller's mu, write mu / logdet / info into the result record.
__global__ void __launch_bounds__(FIN_THREADS)
k_solve_mu(const double* __restrict__ part, int blocks, int p, int mu_mode,
           const double* __restrict__ mu_in, const double* __restrict__ logdet_part, int nt,
           const int* __restrict__ info, double* __restrict__ res, double* __restrict__ gram_out,
           const int* __restrict__ job_slot) {
  extern __shared__ double sm[];
  const int m = p + 1;
  const int npairs = m * (m + 1) / 2;
  {                                                          // job blockIdx.x of a batched launch
    const int j = blockIdx.x, slot = job_slot ? job_slot[j] : 0;
    part += (int64_t)j * blocks * 2 * npairs;
    if (mu_in) mu_in += (int64_t)j * 128;
    logdet_part += (int64_t)slot * nt;
    info += slot;
    res += (int64_t)j * VC_RES_STRIDE;
    // the Gram matrix kept for vc_post is the one of slot 0 = the last evaluation (first job that reads slot 0)
    if (slot != 0) gram_out = nullptr;
  }
  double* G = sm;                 // m x m (column-major, symmetric)
  double* mu = G + m * m;         // p
  const int tid = threadIdx.x;
  for (int pi = tid; pi < npairs; pi += FIN_THREADS) {
    double hi = 0.0, lo = 0.0;
    for (int b = 0; b < blocks; ++b)
      dd_add(hi, lo, part[(int64_t)b * 2 * npairs + 2 * pi], part[(int64_t)b * 2 * npairs + 2 * pi + 1]);
    int rb = (int)((sqrt(8.0 * pi + 1.0) - 1.0) * 0.5);
    while ((rb + 1) * (rb + 2) / 2 <= pi) ++rb;
    while (rb * (rb + 1) / 2 > pi) --rb;
    const int ra = pi - rb * (rb + 1) / 2;
    const double v = hi + lo;
    G[ra + m * rb] = v;
    G[rb + m * ra] = v;
  }
  __syncthreads();
  if (gram_out)
    for (int i = tid; i < m * m; i += FIN_THREADS) gram_out[i] = G[i];
  __syncthreads();
  if (tid == 0) {
    if (mu_mode == VC_MU_GLS) {
      // Cholesky of G_XX (p x p) in place (lower), then two triangular solves with G_Xy
      bool ok = true;
      for (int j = 0; j < p && ok; ++j) {
        double s = G[j + m * j];
        for (int k = 0; k < j; ++k) s -= G[j + m * k] * G[j + m * k];
        if (!(s > 0.0)) { ok = false; break; }
        const double dj = sqrt(s);
        G[j + m * j] = dj;
        for (int i = j + 1; i < p; ++i) {
          double t = G[i + m * j];
          for (int k = 0; k < j; ++k) t -= G[i + m * k] * G[j + m * k];
          G[i + m * j] = t / dj;
        }
      }
      if (ok) {
        for (int i = 0; i < p; ++i) {      // forward: L w = G_Xy  (G_Xy sits in row p: G[p + m i])
          double t = G[p + m * i];
          for (int k = 0; k < i; ++k) t -= G[i + m * k] * mu[k];
          mu[i] = t / G[i + m * i];
        }
        for (int i = p - 1; i >= 0; --i) { // backward: L^T mu = w
          double t = mu[i];
          for (int k = i + 1; k < p; ++k) t -= G[k + m * i] * mu[k];
          mu[i] = t / G[i + m * i];
        }
      } else {
        for (int i = 0; i < p; ++i) mu[i] = nan("");
      }
    } else {
      for (int i = 0; i < p; ++i) mu[i] = mu_in[i];
    }
    for (int i = 0; i < p; ++i) res[VC_RES_MU + i] = mu[i];
    double ld = 0.0, lc = 0.0;             // Kahan over the per-tile partial log-dets, tile order
    for (int t = 0; t < nt; ++t) {
      const double yv = logdet_part[t] - lc;
      const double tv = ld + yv;
      lc = (tv - ld) - yv;
      ld = tv;
    }
    res[VC_RES_LOGDET] = 2.0 * ld;
    const int inf = *info;
    res[VC_RES_INFO] = (inf == INT_MAX) ? 0.0 : (double)inf;
  }
}

// partial sums of (Z_y - Z_X mu)^2 over a slice of columns
__global__ void __launch_bounds__(FIN_THREADS)
k_quad_partial(const double* __restrict__ b, int64_t ldb, int64_t n, int p,
               const double* __restrict__ res, double* __restrict__ part, int64_t b_es,
               const int* __restrict__ job_slot) {
  __shared__ double smu[128];
  __shared__ double shi[FIN_THREADS], slo[FIN_THREADS];
  const int tid = threadIdx.x;
  b += (job_slot ? job_slot[blockIdx.y] : 0) * b_es;        // job blockIdx.y reads the border of its slot
  res += (int64_t)blockIdx.y * VC_RES_STRIDE;
  part += (int64_t)blockIdx.y * gridDim.x * 2;
  if (tid < p) smu[tid] = res[VC_RES_MU + tid];
  __syncthreads();
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t j0 = (int64_t)blockIdx.x * per;
  int64_t j1 = j0 + per;
  if (j1 > n) j1 = n;
  double hi = 0.0, lo = 0.0;
  for (int64_t j = j0 + tid; j < j1; j += FIN_THREADS) {
    const double* zc = b + j * ldb;
    double r = zc[p];
    for (int c = 0; c < p; ++c) r = fma(-zc[c], smu[c], r);
    dd_fma_acc(hi, lo, r, r);
  }
  shi[tid] = hi;
  slo[tid] = lo;
  __syncthreads();
  for (int s = FIN_THREADS / 2; s > 0; s >>= 1) {
    if (tid < s) {
      double h = shi[tid], l = slo[tid];
      dd_add(h, l, shi[tid + s], slo[tid + s]);
      shi[tid] = h;
      slo[tid] = l;
    }
    __syncthreads();
  }
  if (tid == 0) { part[2 * blockIdx.x] = shi[0]; part[2 * blockIdx.x + 1] = slo[0]; }
}

__global__ void k_assemble(const double* __restrict__ part, int blocks, int64_t n, double* __restrict__ res) {
  part += (int64_t)blockIdx.x * blocks * 2;                 // job of a batched launch
  res += (int64_t)blockIdx.x * VC_RES_STRIDE;
  double hi = 0.0, lo = 0.0;
  for (int b = 0; b < blocks; ++b) dd_add(hi, lo, part[2 * b], part[2 * b + 1]);
  const double quad = hi + lo;
  res[VC_RES_QUAD] = quad;
  // same association as the reference: (n log(2 pi) + 2 modulus) + quad
  const double v = ((double)n * log(2.0 * 3.14159265358979323846) + res[VC_RES_LOGDET]) + quad;
  res[VC_RES_N2LL] = (res[VC_RES_INFO] != 0.0) ? nan("") : v;
}

__global__ void k_gather_whitened(const double* __restrict__ b, int64_t ldb, int64_t n,
                                  int p, double* __restrict__ z) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  for (int c = 0; c <= p; ++c) z[j + (int64_t)c * n] = b[c + j * ldb];
}

// R = L^T as base::chol returns it: upper triangular n x n, zeros below the diagonal
__global__ void k_extract_chol(double* a, const VcColBlock* __restrict__ cb, int blk, int64_t n, double* __restrict__ r) {
  __shared__ double t[32][33];
  const int64_t bi = (int64_t)blockIdx.x * 32, bj = (int64_t)blockIdx.y * 32;   // block of L: rows bi.., cols bj..
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int yy = ty; yy < 32; yy += 8) {
    const int64_t i = bi + tx, j = bj + yy;
    t[yy][tx] = (i < n && j < n && i >= j) ? *vc_elem_ptr(a, cb, blk, i, j) : 0.0;
  }
  __syncthreads();
  for (int yy = ty; yy < 32; yy += 8) {
    // R(j, i) = L(i, j): write rows bj.. (fast index), column bi + yy
    const int64_t rj = bj + tx, ci = bi + yy;
    if (rj < n && ci < n) r[rj + ci * n] = t[tx][yy];
  }
}

// out[2i] = sum over blocks of (hi + lo) of pair i, out[2i+1] = 0 (fixed order); distributed finalize
__global__ void k_reduce_dd(const double* __restrict__ part, int blocks, int count, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double hi = 0.0, lo = 0.0;
  for (int b = 0; b < blocks; ++b) dd_add(hi, lo, part[(int64_t)b * 2 * count + 2 * i], part[(int64_t)b * 2 * count + 2 * i + 1]);
  out[2 * i] = hi + lo;
  out[2 * i + 1] = 0.0;
}
__global__ void k_sum_serial(const double* __restrict__ v, int n, double* __restrict__ out) {
  double s = 0.0, c = 0.0;
  for (int i = 0; i < n; ++i) {
    const double yv = v[i] - c;
    const double t = s + yv;
    c = (t - s) - yv;
    s = t;
  }
  *out = s;
}

}  // namespace

// pieces of the finalize for the distributed driver (an NCCL all-reduce goes between them)
void vc_launch_gram_local(const double* b, int64_t ldb, int64_t ncols, int p, VcFinalWork& fw, double* gram_dd,
                          cudaStream_t st, int64_t* launches) {
  const int m = p + 1, npairs = m * (m + 1) / 2;
  const size_t sm1 = ((size_t)m * CHUNK + 2 * (size_t)npairs) * sizeof(double);
  cudaFuncSetAttribute(k_gram_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_gram_partial<<<fw.blocks, FIN_THREADS, sm1, st>>>(b, ldb, ncols, p, fw.gram_part, 0, nullptr);
  k_reduce_dd<<<(npairs + 127) / 128, 128, 0, st>>>(fw.gram_part, fw.blocks, npairs, gram_dd);
  *launches += 2;
}
void vc_launch_sum(const double* v, int n, double* out, cudaStream_t st, int64_t* launches) {
  k_sum_serial<<<1, 1, 0, st>>>(v, n, out);
  ++*launches;
}
void vc_launch_solve_reduced(const double* gram_dd, int p, int mu_mode, const double* mu_in_dev,
                             const double* logdet_sum, const int* info, double* result_dev, double* gram_out,
                             cudaStream_t st, int64_t* launches) {
  const int m = p + 1;
  const size_t sm2 = ((size_t)m * m + p + 1) * sizeof(double);
  cudaFuncSetAttribute(k_solve_mu, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_solve_mu<<<1, FIN_THREADS, sm2, st>>>(gram_dd, 1, p, mu_mode, mu_in_dev, logdet_sum, 1, info, result_dev, gram_out, nullptr);
  ++*launches;
}
void vc_launch_quad_local(const double* b, int64_t ldb, int64_t ncols, int p, const double* result_dev,
                          VcFinalWork& fw, double* quad_dd, cudaStream_t st, int64_t* launches) {
  k_quad_partial<<<fw.blocks, FIN_THREADS, 0, st>>>(b, ldb, ncols, p, result_dev, fw.quad_part, 0, nullptr);
  k_reduce_dd<<<1, 32, 0, st>>>(fw.quad_part, fw.blocks, 1, quad_dd);
  *launches += 2;
}
void vc_launch_assemble_reduced(const double* quad_dd, int64_t n, double* result_dev, cudaStream_t st,
                                int64_t* launches) {
  k_assemble<<<1, 1, 0, st>>>(quad_dd, 1, n, result_dev);
  ++*launches;
}

void vc_launch_finalize(const VcMatrix& M, int p, int mu_mode, const double* mu_in_dev,
                        const double* logdet_part, const int* info, VcFinalWork& fw,
                        double* result_dev, cudaStream_t st, int64_t* launches, double* gram_out_dev,
                        int njobs, const int* job_slot_dev) {
  const int m = p + 1;
  const int npairs = m * (m + 1) / 2;
  const size_t sm1 = ((size_t)m * CHUNK + 2 * (size_t)npairs) * sizeof(double);
  const size_t sm2 = ((size_t)m * m + p + 1) * sizeof(double);
  static bool attr_dev[64] = {false};
  int dev = 0;
  VC_CUDA_TRY(cudaGetDevice(&dev));
  bool& attr = attr_dev[dev & 63];
  if (!attr) {
    VC_CUDA_TRY(cudaFuncSetAttribute(k_gram_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VC_CUDA_TRY(cudaFuncSetAttribute(k_solve_mu, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const unsigned nj = (unsigned)std::max(njobs, 1);
  k_gram_partial<<<dim3(fw.blocks, nj), FIN_THREADS, sm1, st>>>(M.b, M.ldb, M.n, p, fw.gram_part, M.b_es, job_slot_dev);
  k_solve_mu<<<nj, FIN_THREADS, sm2, st>>>(fw.gram_part, fw.blocks, p, mu_mode, mu_in_dev, logdet_part,
                                           M.nt, info, result_dev, gram_out_dev, job_slot_dev);
  k_quad_partial<<<dim3(fw.blocks, nj), FIN_THREADS, 0, st>>>(M.b, M.ldb, M.n, p, result_dev, fw.quad_part, M.b_es, job_slot_dev);
  k_assemble<<<nj, 1, 0, st>>>(fw.quad_part, fw.blocks, M.n, result_dev);
  *launches += 4;
}

void vc_launch_gather_whitened(const VcMatrix& M, int p, double* z_out_dev, cudaStream_t st,
                               int64_t* launches) {
  k_gather_whitened<<<(unsigned)((M.n + 255) / 256), 256, 0, st>>>(M.b, M.ldb, M.n, p, z_out_dev);
  ++*launches;
}

void vc_launch_extract_chol(const VcMatrix& M, double* r_out_dev, cudaStream_t st, int64_t* launches) {
  dim3 grid((unsigned)((M.n + 31) / 32), (unsigned)((M.n + 31) / 32));
  k_extract_chol<<<grid, dim3(32, 8), 0, st>>>(M.a, M.cb, M.blk, M.n, r_out_dev);
  ++*launches;
}
