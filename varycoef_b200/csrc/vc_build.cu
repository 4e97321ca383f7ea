// vc_build.cu -- K1: fused pairwise-distance + covariance build of Sigma_y (lower tiles).
//
// Replaces, in one write-only pass over HBM:
//   own_dist            reference R-pkg/R/utils.R:508-543   (distance matrix, never materialised)
//   cov.func closure    R-pkg/R/SVC_mle.R:58-61 + R-pkg/R/utils.R:2-29 (spam::cov.*)
//   outer.W (x taper)   R-pkg/R/SVC_mle.R:65-72, get_taper R-pkg/R/utils.R:545-552
//   Sigma_y             R-pkg/R/utils.R:171-233 (sum_k cov_k o W_k W_k^T + tau^2 I)
//
// One CTA = one 128x128 tile (I >= J).  Thread t owns rows 2*(t&63), +1 (one 16-byte store per
// column, 512 contiguous bytes per warp) and the 32 columns (t>>6)*32.. of the tile.  Column
// operands (coordinates, W_jk) are broadcast from shared memory; row operands stay in registers
// when d <= 3 and q <= 8.  Padded rows/cols (index >= n) get W = 0 and a unit diagonal so the
// padded matrix factors to [L 0; 0 I].
#include "vc_common.cuh"
#include "vc_math.cuh"
#include <algorithm>

namespace {

constexpr int TILE = VC_TILE;
constexpr int BUILD_THREADS = 256;
constexpr int RD = 3;   // register-resident coordinate columns
constexpr int RQ = 8;   // register-resident SVCs

__constant__ double c_exp2_tab[VC_EXP_TABLE_SIZE];

struct BuildArgs {
  double* out;
  int64_t ld;
  int64_t n, n_pad;
  int nt;
  int d;
  const double* locs;   // n_pad x d
  const double* w;      // n_pad x q
  const int* tiles;     // optional list of packed (I | J << 16) global tiles (distributed build)
  const VcColBlock* cb; // block-column storage of the OUTPUT (lower tiles); nullptr = dense `out` with `ld` (FULL)
  int blk, pr, pc;      // tiles per storage block; process grid of a block-cyclic output (1 x 1 = plain)
  int64_t out_es;       // batched build: slot e = blockIdx.y writes out + e * out_es with the parameters th_dev[e]
  const VcTheta* th_dev;
  VcTheta th;           // model (cov / taper / distance / q) and, when th_dev == nullptr, the parameters
};
// covariance parameters of this CTA's evaluation slot: by-value block (single) or the device array (batched)
__device__ __forceinline__ double th_inv_range(const BuildArgs& p, int k) {
  return p.th_dev ? p.th_dev[blockIdx.y].inv_range[k] : p.th.inv_range[k];
}
__device__ __forceinline__ double th_sill(const BuildArgs& p, int k) {
  return p.th_dev ? p.th_dev[blockIdx.y].sill[k] : p.th.sill[k];
}
__device__ __forceinline__ double th_tau2(const BuildArgs& p) {
  return p.th_dev ? p.th_dev[blockIdx.y].tau2 : p.th.tau2;
}

// shared layout: tab[64] | colX[d][128] | colW[q][128] | rowX[d][128] | rowA[q][128] (rowA = W rows)
template <int COV, bool REG, bool FULL>
__global__ void __launch_bounds__(BUILD_THREADS) k_build(BuildArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int d = p.d, q = p.th.q;
  double* tab = sm;
  double* colX = tab + VC_EXP_TABLE_SIZE;
  double* colW = colX + d * TILE;
  double* rowX = colW + q * TILE;
  double* rowA = rowX + d * TILE;

  int I, J;
  if (p.tiles) {
    const int packed = p.tiles[blockIdx.x];
    I = packed & 0xffff;
    J = packed >> 16;
  } else if (FULL) {
    I = blockIdx.x % p.nt;
    J = blockIdx.x / p.nt;
  } else {
    // lower-triangular tile enumeration: id = I (I + 1) / 2 + J
    const int id = blockIdx.x;
    I = (int)((sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= id) ++I;
    while (I * (I + 1) / 2 > id) --I;
    J = id - I * (I + 1) / 2;
  }
  const int tid = threadIdx.x;
  if (tid < VC_EXP_TABLE_SIZE) tab[tid] = c_exp2_tab[tid];
  for (int idx = tid; idx < d * TILE; idx += BUILD_THREADS) {
    const int c = idx / TILE, r = idx % TILE;
    colX[idx] = p.locs[(int64_t)c * p.n_pad + (int64_t)J * TILE + r];
    rowX[idx] = p.locs[(int64_t)c * p.n_pad + (int64_t)I * TILE + r];
  }
  for (int idx = tid; idx < q * TILE; idx += BUILD_THREADS) {
    const int k = idx / TILE, r = idx % TILE;
    colW[idx] = p.w[(int64_t)k * p.n_pad + (int64_t)J * TILE + r];
    rowA[idx] = p.w[(int64_t)k * p.n_pad + (int64_t)I * TILE + r];
  }
  __syncthreads();

  const int rg = tid & 63, cg = tid >> 6;
  const int r0 = 2 * rg;
  const int64_t gi0 = (int64_t)I * TILE + r0;
  // position in the (possibly block-cyclic local) block-column storage
  double* otile = nullptr;
  int64_t old_ = 0;
  if (!FULL) {
    const int lI = p.pr > 1 ? ((I / p.blk) / p.pr) * p.blk + I % p.blk : I;
    const int lJ = p.pc > 1 ? ((J / p.blk) / p.pc) * p.blk + J % p.blk : J;
    otile = vc_tile_ptr(p.out + blockIdx.y * p.out_es, p.cb, p.blk, lI, lJ, old_) + r0;
  }
  const double tau2 = th_tau2(p);
  const int method = p.th.dist_method;
  const double mink = p.th.mink_p;

  double xi0[RD], xi1[RD], ai0[RQ], ai1[RQ], ir[RQ], sk[RQ];
  if (REG) {
#pragma unroll
    for (int c = 0; c < RD; ++c)
      if (c < d) { xi0[c] = rowX[c * TILE + r0]; xi1[c] = rowX[c * TILE + r0 + 1]; }
#pragma unroll
    for (int k = 0; k < RQ; ++k)
      if (k < q) { ai0[k] = rowA[k * TILE + r0]; ai1[k] = rowA[k * TILE + r0 + 1]; ir[k] = th_inv_range(p, k); sk[k] = th_sill(p, k); }
  }

  for (int cc = 0; cc < 32; ++cc) {
    const int j = cg * 32 + cc;
    const int64_t gj = (int64_t)J * TILE + j;
    double s0 = 0.0, s1 = 0.0;
    if (REG) {
#pragma unroll
      for (int c = 0; c < RD; ++c)
        if (c < d) {
          const double xj = colX[c * TILE + j];
          if (method == VC_DIST_EUCLIDEAN) {
            const double d0 = xi0[c] - xj, d1 = xi1[c] - xj;
            s0 = fma(d0, d0, s0);
            s1 = fma(d1, d1, s1);
          } else {
            s0 = vc_dist_accum(method, s0, xi0[c] - xj, mink);
            s1 = vc_dist_accum(method, s1, xi1[c] - xj, mink);
          }
        }
    } else {
      for (int c = 0; c < d; ++c) {
        const double xj = colX[c * TILE + j];
        s0 = vc_dist_accum(method, s0, rowX[c * TILE + r0] - xj, mink);
        s1 = vc_dist_accum(method, s1, rowX[c * TILE + r0 + 1] - xj, mink);
      }
    }
    double h0, h1;
    if (method == VC_DIST_EUCLIDEAN) { h0 = sqrt(s0); h1 = sqrt(s1); }
    else { h0 = vc_dist_finish(method, s0, mink); h1 = vc_dist_finish(method, s1, mink); }

    double v0 = 0.0, v1 = 0.0;
    if (REG) {
#pragma unroll
      for (int k = 0; k < RQ; ++k)
        if (k < q) {
          const double wj = colW[k * TILE + j];
          // (W_ik W_jk) * (sill_k * shape): the reference's Cov * outer.W, exactly symmetric in (i, j)
          v0 = fma(ai0[k] * wj, sk[k] * vc_cov_shape<COV>(h0 * ir[k], tab), v0);
          v1 = fma(ai1[k] * wj, sk[k] * vc_cov_shape<COV>(h1 * ir[k], tab), v1);
        }
    } else {
      for (int k = 0; k < q; ++k) {
        const double wj = colW[k * TILE + j];
        const double irk = th_inv_range(p, k);
        const double sk_ = th_sill(p, k);
        v0 = fma(rowA[k * TILE + r0] * wj, sk_ * vc_cov_shape<COV>(h0 * irk, tab), v0);
        v1 = fma(rowA[k * TILE + r0 + 1] * wj, sk_ * vc_cov_shape<COV>(h1 * irk, tab), v1);
      }
    }
    if (p.th.taper_id != VC_TAPER_NONE) {
      v0 *= vc_taper(p.th.taper_id, h0, p.th.inv_delta);
      v1 *= vc_taper(p.th.taper_id, h1, p.th.inv_delta);
    }
    if (gi0 == gj) v0 += (gj < p.n) ? tau2 : 1.0;
    if (gi0 + 1 == gj) v1 += (gj < p.n) ? tau2 : 1.0;
    if (FULL) {
      if (gj < p.n) {
        if (gi0 < p.n) p.out[gi0 + gj * p.ld] = v0;
        if (gi0 + 1 < p.n) p.out[gi0 + 1 + gj * p.ld] = v1;
      }
    } else {
      *reinterpret_cast<double2*>(otile + (int64_t)j * old_) = make_double2(v0, v1);
    }
  }
}


// Fast path of K1: Euclidean distance, D <= 3 coordinate columns, Q <= 4 SVCs, everything a
// compile-time constant (no per-term predicates, row operands in registers, two columns in flight
// per thread for instruction-level parallelism).  Same arithmetic as k_build.
template <int COV, int D, int Q>
__global__ void __launch_bounds__(BUILD_THREADS, 3) k_build_fast(BuildArgs p) {
  __shared__ __align__(16) double tab[VC_EXP_TABLE_SIZE];
  __shared__ __align__(16) double colX[D][TILE];
  __shared__ __align__(16) double colW[Q][TILE];
  int I, J;
  if (p.tiles) {
    const int packed = p.tiles[blockIdx.x];
    I = packed & 0xffff;
    J = packed >> 16;
  } else {
    const int id = blockIdx.x;
    I = (int)((sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= id) ++I;
    while (I * (I + 1) / 2 > id) --I;
    J = id - I * (I + 1) / 2;
  }
  const int tid = threadIdx.x;
  if (tid < VC_EXP_TABLE_SIZE) tab[tid] = c_exp2_tab[tid];
  for (int idx = tid; idx < D * TILE; idx += BUILD_THREADS)
    colX[idx / TILE][idx % TILE] = p.locs[(int64_t)(idx / TILE) * p.n_pad + (int64_t)J * TILE + idx % TILE];
  for (int idx = tid; idx < Q * TILE; idx += BUILD_THREADS)
    colW[idx / TILE][idx % TILE] = p.w[(int64_t)(idx / TILE) * p.n_pad + (int64_t)J * TILE + idx % TILE];
  const int rg = tid & 63, cg = tid >> 6;
  const int r0 = 2 * rg;
  const int64_t gi0 = (int64_t)I * TILE + r0;
  double xi0[D], xi1[D], wi0[Q], wi1[Q], ir[Q], sk[Q];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const double2 v = *reinterpret_cast<const double2*>(p.locs + (int64_t)c * p.n_pad + gi0);
    xi0[c] = v.x; xi1[c] = v.y;
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    const double2 v = *reinterpret_cast<const double2*>(p.w + (int64_t)k * p.n_pad + gi0);
    wi0[k] = v.x; wi1[k] = v.y;
    ir[k] = th_inv_range(p, k);
    sk[k] = th_sill(p, k);
  }
  const double tau2 = th_tau2(p);
  const int lI = p.pr > 1 ? ((I / p.blk) / p.pr) * p.blk + I % p.blk : I;
  const int lJ = p.pc > 1 ? ((J / p.blk) / p.pc) * p.blk + J % p.blk : J;
  int64_t old_;
  double* out = vc_tile_ptr(p.out + blockIdx.y * p.out_es, p.cb, p.blk, lI, lJ, old_) + r0;
  const int taper_id = p.th.taper_id;
  const double inv_delta = p.th.inv_delta;
  __syncthreads();
#pragma unroll 2
  for (int cc = 0; cc < 32; ++cc) {
    const int j = cg * 32 + cc;
    const int64_t gj = (int64_t)J * TILE + j;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const double xj = colX[c][j];
      const double d0 = xi0[c] - xj, d1 = xi1[c] - xj;
      s0 = fma(d0, d0, s0);
      s1 = fma(d1, d1, s1);
    }
    const double h0 = vc_sqrt_pos(s0), h1 = vc_sqrt_pos(s1);
    double v0 = 0.0, v1 = 0.0;
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      const double wj = colW[k][j];
      v0 = fma(wi0[k] * wj, sk[k] * vc_cov_shape<COV>(h0 * ir[k], tab), v0);
      v1 = fma(wi1[k] * wj, sk[k] * vc_cov_shape<COV>(h1 * ir[k], tab), v1);
    }
    if (taper_id != VC_TAPER_NONE) {
      v0 *= vc_taper(taper_id, h0, inv_delta);
      v1 *= vc_taper(taper_id, h1, inv_delta);
    }
    if (gi0 == gj) v0 += (gj < p.n) ? tau2 : 1.0;
    if (gi0 + 1 == gj) v1 += (gj < p.n) ? tau2 : 1.0;
    *reinterpret_cast<double2*>(out + (int64_t)j * old_) = make_double2(v0, v1);
  }
}

template <int COV, int D>
bool launch_fast_q(const BuildArgs& a, dim3 grid, cudaStream_t st) {
  switch (a.th.q) {
    case 1: k_build_fast<COV, D, 1><<<grid, BUILD_THREADS, 0, st>>>(a); return true;
    case 2: k_build_fast<COV, D, 2><<<grid, BUILD_THREADS, 0, st>>>(a); return true;
    case 3: k_build_fast<COV, D, 3><<<grid, BUILD_THREADS, 0, st>>>(a); return true;
    case 4: k_build_fast<COV, D, 4><<<grid, BUILD_THREADS, 0, st>>>(a); return true;
    default: return false;
  }
}
template <int COV>
bool launch_fast_d(const BuildArgs& a, dim3 grid, cudaStream_t st) {
  switch (a.d) {
    case 1: return launch_fast_q<COV, 1>(a, grid, st);
    case 2: return launch_fast_q<COV, 2>(a, grid, st);
    case 3: return launch_fast_q<COV, 3>(a, grid, st);
    default: return false;
  }
}
// exp / Matern with Euclidean distances, d <= 3, q <= 4: the common case of the reference's API
bool launch_fast(const BuildArgs& a, dim3 grid, cudaStream_t st) {
  if (a.th.dist_method != VC_DIST_EUCLIDEAN) return false;
  switch (a.th.cov_id) {
    case VC_COV_EXP: return launch_fast_d<0>(a, grid, st);
    case VC_COV_MAT32: return launch_fast_d<1>(a, grid, st);
    case VC_COV_MAT52: return launch_fast_d<2>(a, grid, st);
    default: return false;
  }
}

template <bool FULL>
void launch_build_t(const BuildArgs& a, cudaStream_t st, unsigned grid_override = 0, int nbatch = 1) {
  const int d = a.d, q = a.th.q;
  const bool reg = (d <= RD && q <= RQ);
  const size_t smem = (VC_EXP_TABLE_SIZE + 2 * (size_t)(d + q) * TILE) * sizeof(double);
  const dim3 grid(grid_override ? grid_override
                  : FULL ? (unsigned)a.nt * a.nt : (unsigned)((int64_t)a.nt * (a.nt + 1) / 2), (unsigned)std::max(nbatch, 1));
  if (!FULL && launch_fast(a, grid, st)) return;
#define VC_BUILD_CASE(C)                                                               \
  case C:                                                                              \
    if (reg) k_build<C, true, FULL><<<grid, BUILD_THREADS, smem, st>>>(a);             \
    else k_build<C, false, FULL><<<grid, BUILD_THREADS, smem, st>>>(a);                \
    break;
  switch (a.th.cov_id) {
    VC_BUILD_CASE(0) VC_BUILD_CASE(1) VC_BUILD_CASE(2) VC_BUILD_CASE(3) VC_BUILD_CASE(4)
    default:
      if (reg) k_build<5, true, FULL><<<grid, BUILD_THREADS, smem, st>>>(a);
      else k_build<5, false, FULL><<<grid, BUILD_THREADS, smem, st>>>(a);
  }
#undef VC_BUILD_CASE
}

// border tile 0: B^T[c, j] = X[j, c] (c < p), y[j] (c == p), 0 otherwise
__global__ void k_border(double* b, int64_t ldb, int64_t n_pad, int p, const double* __restrict__ x,
                         const double* __restrict__ y, int64_t b_es) {
  b += blockIdx.y * b_es;
  const int64_t j = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7);
  const int c = threadIdx.x & 127;
  if (j >= n_pad) return;
  double v = 0.0;
  if (c < p) v = x[(int64_t)c * n_pad + j];
  else if (c == p) v = y[j];
  b[c + j * ldb] = v;
}

}  // namespace

void vc_build_init() {
  static bool done_dev[64] = {false};
  int dev = 0;
  VC_CUDA_TRY(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return;
  double tab[VC_EXP_TABLE_SIZE];
  vc_fill_exp2_table(tab);
  VC_CUDA_TRY(cudaMemcpyToSymbol(c_exp2_tab, tab, sizeof(tab)));
  done = true;
}

void vc_launch_build(const VcMatrix& M, const VcTheta& th, int d, const double* locs_pad,
                     const double* w_pad, cudaStream_t st, int64_t* launches, const VcTheta* th_dev) {
  BuildArgs a{};
  a.out = M.a; a.cb = M.cb; a.blk = M.blk; a.pr = 1; a.pc = 1; a.n = M.n; a.n_pad = M.n_pad; a.nt = M.nt; a.d = d;
  a.locs = locs_pad; a.w = w_pad; a.th = th;
  a.th_dev = th_dev; a.out_es = M.a_es;
  launch_build_t<false>(a, st, 0, th_dev ? M.nbatch : 1);
  ++*launches;
}

// distributed build: only the listed global tiles, written at their block-cyclic local position
void vc_launch_build_tiles(const VcMatrix& M, const VcTheta& th, int d,
                           const double* locs_pad, const double* w_pad, const int* tiles, int ntiles,
                           int pr, int pc, cudaStream_t st, int64_t* launches) {
  if (ntiles <= 0) return;
  BuildArgs a{};
  a.out = M.a; a.cb = M.cb; a.n = M.n; a.n_pad = M.n_pad; a.nt = M.nt; a.d = d;
  a.locs = locs_pad; a.w = w_pad; a.th = th;
  a.tiles = tiles; a.blk = M.blk; a.pr = pr; a.pc = pc;
  launch_build_t<false>(a, st, (unsigned)ntiles);
  ++*launches;
}

// border tile 0 of a block-cyclic column distribution: local column tile lj <-> global tile gj[lj]
__global__ void k_border_dist(double* b, int64_t ldb, int64_t n_pad, int p, const double* __restrict__ x,
                              const double* __restrict__ y, const int* __restrict__ col_tiles, int nloc) {
  const int lj = blockIdx.x;
  const int64_t gj0 = (int64_t)col_tiles[lj] * TILE;
  for (int idx = threadIdx.x; idx < TILE * TILE; idx += blockDim.x) {
    const int c = idx & (TILE - 1), jj = idx >> 7;
    const int64_t j = gj0 + jj;
    double v = 0.0;
    if (c < p) v = x[(int64_t)c * n_pad + j];
    else if (c == p) v = y[j];
    b[c + ((int64_t)lj * TILE + jj) * ldb] = v;
  }
}
void vc_launch_border_dist(double* b, int64_t ldb, int64_t n_pad, int p, const double* x_pad,
                           const double* y_pad, const int* col_tiles_dev, int nloc, cudaStream_t st,
                           int64_t* launches) {
  if (nloc <= 0) return;
  k_border_dist<<<nloc, 256, 0, st>>>(b, ldb, n_pad, p, x_pad, y_pad, col_tiles_dev, nloc);
  ++*launches;
}

void vc_launch_sigma_full(double* out, int64_t n, int64_t n_pad, const VcTheta& th, int d,
                          const double* locs_pad, const double* w_pad, cudaStream_t st,
                          int64_t* launches) {
  BuildArgs a{};
  a.out = out; a.ld = n; a.n = n; a.n_pad = n_pad; a.nt = (int)(n_pad / TILE); a.d = d;
  a.locs = locs_pad; a.w = w_pad; a.th = th;
  a.blk = 1; a.pr = 1; a.pc = 1;
  launch_build_t<true>(a, st);
  ++*launches;
}

void vc_launch_border(const VcMatrix& M, int p, const double* x_pad, const double* y_pad,
                      cudaStream_t st, int64_t* launches) {
  k_border<<<dim3((unsigned)((M.n_pad + 1) / 2), (unsigned)std::max(M.nbatch, 1)), 256, 0, st>>>(M.b, M.ldb, M.n_pad, p, x_pad, y_pad, M.b_es);
  ++*launches;
}
