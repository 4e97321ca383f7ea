// vc_chol.cu -- K2: bordered, blocked, right-looking fp64 Cholesky on sm_100a.
//
// Replaces base::chol (LAPACK dpotrf) at reference R-pkg/R/objective_functions.R:26 and, through
// the border rows, the triangular solves solve(t(R), X), solve(t(R), y), solve(t(R), res) of
// R-pkg/R/utils.R:94,100 and R-pkg/R/objective_functions.R:49.
//
// Storage: lower triangle of Sigma in a column-major (n_pad + 128) x n_pad array; the 128 rows
// below the matrix hold B^T = [X y]^T.  Running the factorisation over that bordered matrix turns
// the border into Z^T = (L^-1 B)^T for free (each panel step treats it as one more row tile).
//
// Kernels:
//   k_potrf128  one CTA: factor a 128x128 diagonal tile in shared memory, also its inverse.
//   k_trsm_panel / k_syrk_update  128x128 CTA tile, 8 warps x (64x32) warp tiles, DMMA m8n8k4 (fp64 tensor pipe),
//               4-stage cp.async pipeline.  MODE_TRSM: X = A * Linv^T (panel solve as a GEMM);
//               MODE_UPDATE: C -= P_I * P_J^T (inner panel update and trailing SYRK).
#include "vc_common.cuh"
#include <limits.h>

namespace {

constexpr int TILE = VC_TILE;
constexpr int KC = 16;        // k-depth of one pipeline stage
constexpr int STAGES = 4;
constexpr int SLD = 132;      // shared-memory stride of one k-column: 128 + 4 doubles -> the
                              // (row, k) fragment pattern of m8n8k4 hits 16 distinct bank pairs
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM = 2 * STAGES * KC * SLD * (int)sizeof(double);  // 135168 B
constexpr int SUPER = 16;     // super-block edge (tiles) of the L2-friendly rasterisation

enum { MODE_TRSM = 0, MODE_UPDATE = 1 };
enum { RASTER_RECT = 0, RASTER_SUPER = 1 };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4, row) * B(4x8, col) on the fp64 tensor pipe (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct GemmArgs {
  double* a;        // bordered matrix
  int64_t lda;
  const double* linv;   // MODE_TRSM: 128x128 inverse factor (ld 128)
  int k0;           // first panel column (elements)
  int K;            // panel depth (multiple of KC)
  int jt;           // MODE_TRSM: panel tile column
  int J0, J1;       // MODE_UPDATE: tile columns [J0, J1)
  int nrows;        // number of row tiles incl. border
  int raster;
  int nsbj;         // RASTER_SUPER: super-block columns
};

// acc[i][j][0..1]: rows wm*64 + i*8 + (lane>>2), cols wn*32 + j*8 + 2*(lane&3) + {0,1}
__device__ __forceinline__ void gemm_mainloop(const double* __restrict__ Ag, int64_t lda_a,
                                              const double* __restrict__ Bg, int64_t lda_b, int K,
                                              double (&acc)[8][4][2], double* smem) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, tig = lane & 3;
  double* sA = smem;
  double* sB = smem + STAGES * KC * SLD;
  const int nk = K / KC;

  auto load_stage = [&](int slot, int kt) {
    double* sa = sA + slot * KC * SLD;
    double* sb = sB + slot * KC * SLD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = tid + GEMM_THREADS * i;   // 0..1023
      const int col = id >> 6;                 // k column within the stage
      const int rc = (id & 63) * 2;            // row of the 16-byte chunk
      cp_async16(sa + col * SLD + rc, Ag + (int64_t)(kt * KC + col) * lda_a + rc);
      cp_async16(sb + col * SLD + rc, Bg + (int64_t)(kt * KC + col) * lda_b + rc);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < nk) load_stage(nxt % STAGES, nxt);
    cp_async_commit();
    const int slot = kt % STAGES;
    const double* sa = sA + slot * KC * SLD + wm * 64 + g;
    const double* sb = sB + slot * KC * SLD + wn * 32 + g;
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; ++k4) {
      const int kk = (k4 * 4 + tig) * SLD;
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = sa[kk + i * 8];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sb[kk + j * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

template <int MODE>
__device__ __forceinline__ void gemm_body(const GemmArgs& p) {
  extern __shared__ __align__(16) double smem[];
  int I, J;
  if (MODE == MODE_TRSM) {
    I = p.jt + 1 + blockIdx.x;
    J = p.jt;
  } else if (p.raster == RASTER_RECT) {
    const int nr = p.nrows - p.J0;
    J = p.J0 + blockIdx.x / nr;
    I = p.J0 + blockIdx.x % nr;
    if (I < J) return;
  } else {
    // super-blocks of SUPER x SUPER tiles, enumerated row by row over the lower trapezoid
    const int s = blockIdx.x / (SUPER * SUPER);
    const int w = blockIdx.x % (SUPER * SUPER);
    const int T = p.nsbj * (p.nsbj + 1) / 2;
    int sbi, sbj;
    if (s < T) {
      sbi = (int)((sqrt(8.0 * s + 1.0) - 1.0) * 0.5);
      while ((sbi + 1) * (sbi + 2) / 2 <= s) ++sbi;
      while (sbi * (sbi + 1) / 2 > s) --sbi;
      sbj = s - sbi * (sbi + 1) / 2;
    } else {
      sbi = p.nsbj + (s - T) / p.nsbj;
      sbj = (s - T) % p.nsbj;
    }
    I = p.J0 + sbi * SUPER + (w % SUPER);
    J = p.J0 + sbj * SUPER + (w / SUPER);
    if (J >= p.J1 || I >= p.nrows || I < J) return;
  }

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, tig = lane & 3;

  double acc[8][4][2];
  double* Ct = p.a + (int64_t)I * TILE + (int64_t)J * TILE * p.lda;
  double* cbase = Ct + (wm * 64 + g) + (int64_t)(wn * 32 + 2 * tig) * p.lda;

  if (MODE == MODE_UPDATE) {
    // acc starts at -C so that the epilogue is a pure store: C_new = -(-C + sum a b)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double* c = cbase + i * 8 + (int64_t)(j * 8) * p.lda;
        acc[i][j][0] = -c[0];
        acc[i][j][1] = -c[p.lda];
      }
    const double* Pi = p.a + (int64_t)I * TILE + (int64_t)p.k0 * p.lda;
    const double* Pj = p.a + (int64_t)J * TILE + (int64_t)p.k0 * p.lda;
    gemm_mainloop(Pi, p.lda, Pj, p.lda, p.K, acc, smem);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double* c = cbase + i * 8 + (int64_t)(j * 8) * p.lda;
        c[0] = -acc[i][j][0];
        c[p.lda] = -acc[i][j][1];
      }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // X(I, jt) = A(I, jt) * Linv^T : the B operand element (n, k) is Linv[n + 128 k]
    gemm_mainloop(Ct, p.lda, p.linv, TILE, TILE, acc, smem);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double* c = cbase + i * 8 + (int64_t)(j * 8) * p.lda;
        c[0] = acc[i][j][0];
        c[p.lda] = acc[i][j][1];
      }
  }
}

// the two entry points of the DMMA tile kernel (distinct names so profiles read clearly)
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_trsm_panel(GemmArgs p) { gemm_body<MODE_TRSM>(p); }
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_syrk_update(GemmArgs p) { gemm_body<MODE_UPDATE>(p); }

// ---------------------------------------------------------------------------------------------
// k_potrf128: one CTA factors the 128x128 diagonal tile (unblocked right-looking in shared
// memory) and inverts the factor.  Shared array M is 128 x 129 column-major (ld 128):
//   L(i, j)    (i >= j) lives at M[i + 128 j]
//   Linv(i, c) (i >= c) lives at M[c + 128 (i + 1)]   (transposed, shifted one column right)
// so factor and inverse share one 129-KB buffer.
// ---------------------------------------------------------------------------------------------
constexpr int POTRF_THREADS = 256;
constexpr int POTRF_SMEM = (TILE * (TILE + 1) + TILE) * (int)sizeof(double);

__global__ void __launch_bounds__(POTRF_THREADS, 1)
k_potrf128(double* __restrict__ At, int64_t lda, int jt, double* __restrict__ linv_all,
           double* __restrict__ logdet_part, int* __restrict__ info) {
  extern __shared__ __align__(16) double M[];
  double* sdiag = M + TILE * (TILE + 1);
  const int tid = threadIdx.x;
  const int lane = tid & 31, grp = tid >> 5;   // 8 groups of 32

  for (int idx = tid; idx < TILE * TILE; idx += POTRF_THREADS) {
    const int r = idx & (TILE - 1), c = idx >> 7;
    if (r >= c) M[r + TILE * c] = At[r + (int64_t)c * lda];
    // inverse starts as the identity
    if (r >= c) M[c + TILE * (r + 1)] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();

  // ---- factor ----
  for (int j = 0; j < TILE; ++j) {
    const double ajj = M[j + TILE * j];
    const double d = sqrt(ajj);
    if (tid == j) {
      sdiag[j] = d;
      if (!(ajj > 0.0)) atomicMin(info, jt * TILE + j + 1);
    }
    if (tid > j && tid < TILE) M[tid + TILE * j] = M[tid + TILE * j] / d;
    __syncthreads();
    // trailing rank-1 update: M(i, k) -= M(i, j) M(k, j), j < k <= i
    double mij[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) mij[a] = M[lane + 32 * a + TILE * j];
    for (int k = j + 1 + grp; k < TILE; k += 8) {
      const double mkj = M[k + TILE * j];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = lane + 32 * a;
        if (i >= k) M[i + TILE * k] = fma(-mij[a], mkj, M[i + TILE * k]);
      }
    }
    __syncthreads();
  }

  // ---- inverse: rows of X = L^-1 one after the other (right-looking forward substitution) ----
  for (int i = 0; i < TILE; ++i) {
    const double dinv = 1.0 / sdiag[i];
    if (tid <= i) M[tid + TILE * (i + 1)] *= dinv;     // X(i, c), c <= i
    __syncthreads();
    // X(i2, c) -= L(i2, i) X(i, c), i2 > i, c <= i
    for (int i2 = i + 1 + grp; i2 < TILE; i2 += 8) {
      const double l = M[i2 + TILE * i];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int c = lane + 32 * a;
        if (c <= i) M[c + TILE * (i2 + 1)] = fma(-l, M[c + TILE * (i + 1)], M[c + TILE * (i2 + 1)]);
      }
    }
    __syncthreads();
  }

  // ---- write back ----
  double* linv = linv_all + (int64_t)jt * TILE * TILE;
  for (int idx = tid; idx < TILE * TILE; idx += POTRF_THREADS) {
    const int r = idx & (TILE - 1), c = idx >> 7;
    if (r > c) At[r + (int64_t)c * lda] = M[r + TILE * c];
    else if (r == c) At[r + (int64_t)c * lda] = sdiag[r];
    linv[idx] = (r >= c) ? M[c + TILE * (r + 1)] : 0.0;
  }
  if (tid < 32) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) s += log(sdiag[lane * 4 + a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) logdet_part[jt] = s;
  }
}

__global__ void k_reset_info(int* info) { *info = INT_MAX; }

}  // namespace

void vc_chol_init() {
  static bool done_dev[64] = {false};
  int dev = 0;
  VC_CUDA_TRY(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return;
  VC_CUDA_TRY(cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_syrk_update, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_potrf128, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
  done = true;
}

static void launch_update(const VcMatrix& M, int k0, int K, int J0, int J1, cudaStream_t st,
                          int64_t* launches) {
  if (J1 <= J0) return;
  GemmArgs g{};
  g.a = M.a; g.lda = M.lda; g.linv = nullptr; g.k0 = k0; g.K = K; g.jt = 0;
  g.J0 = J0; g.J1 = J1; g.nrows = M.nt + 1;
  const int nj = J1 - J0;
  unsigned grid;
  if (nj < SUPER) {
    g.raster = RASTER_RECT; g.nsbj = 0;
    grid = (unsigned)nj * (unsigned)(g.nrows - J0);
  } else {
    g.raster = RASTER_SUPER;
    g.nsbj = (nj + SUPER - 1) / SUPER;
    const int nsbi = (g.nrows - J0 + SUPER - 1) / SUPER;
    const int T = g.nsbj * (g.nsbj + 1) / 2;
    const int nsb = T + (nsbi - g.nsbj) * g.nsbj;
    grid = (unsigned)nsb * SUPER * SUPER;
  }
  k_syrk_update<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
  ++*launches;
}

// Factor the bordered matrix.  Two-level blocking: 128-wide sub-panels (potrf128 + TRSM-as-GEMM +
// inner update of the rest of the outer block, K = 128) inside nb-wide outer blocks whose
// trailing SYRK runs with K = nb.  With lookahead the next outer block's columns are updated
// first and its panel is factored on a second stream while the rest of the trailing matrix
// is still being updated.
void vc_chol_factor(const VcMatrix& M, VcCholWork& w, int64_t* launches) {
  const int nt = M.nt, nrows = nt + 1, nb = w.nb_tiles;
  cudaStream_t sm = w.st_main, sp = w.lookahead ? w.st_panel : w.st_main;
  k_reset_info<<<1, 1, 0, sm>>>(w.info);
  ++*launches;
  if (w.lookahead) {
    VC_CUDA_TRY(cudaEventRecord(w.ev_update, sm));
    VC_CUDA_TRY(cudaStreamWaitEvent(sp, w.ev_update, 0));
  }
  for (int ob = 0; ob < nt; ob += nb) {
    const int oe = (ob + nb < nt) ? ob + nb : nt;
    // ---- panel: columns [ob, oe) on the panel stream ----
    for (int jt = ob; jt < oe; ++jt) {
      double* diag = M.a + (int64_t)jt * TILE * (M.lda + 1);
      k_potrf128<<<1, POTRF_THREADS, POTRF_SMEM, sp>>>(diag, M.lda, jt, w.linv, w.logdet_part, w.info);
      ++*launches;
      GemmArgs g{};
      g.a = M.a; g.lda = M.lda; g.linv = w.linv + (int64_t)jt * TILE * TILE;
      g.k0 = jt * TILE; g.K = TILE; g.jt = jt; g.nrows = nrows;
      k_trsm_panel<<<nrows - (jt + 1), GEMM_THREADS, GEMM_SMEM, sp>>>(g);
      ++*launches;
      launch_update(M, jt * TILE, TILE, jt + 1, oe, sp, launches);
    }
    if (oe >= nt) break;
    const int K = (oe - ob) * TILE;
    if (!w.lookahead) {
      launch_update(M, ob * TILE, K, oe, nt, sm, launches);
    } else {
      // panel [ob, oe) is final on sp; main stream waits for it
      VC_CUDA_TRY(cudaEventRecord(w.ev_panel, sp));
      VC_CUDA_TRY(cudaStreamWaitEvent(sm, w.ev_panel, 0));
      const int ne = (oe + nb < nt) ? oe + nb : nt;
      // next panel's columns first ...
      launch_update(M, ob * TILE, K, oe, ne, sm, launches);
      VC_CUDA_TRY(cudaEventRecord(w.ev_col, sm));
      // ... then the rest of the trailing matrix, overlapping the next panel on sp
      launch_update(M, ob * TILE, K, ne, nt, sm, launches);
      VC_CUDA_TRY(cudaEventRecord(w.ev_update, sm));
      VC_CUDA_TRY(cudaStreamWaitEvent(sp, w.ev_col, 0));
      // the next panel's inner updates (and TRSM of lower row tiles) only touch columns
      // [oe, ne), which the "rest" update does not write; but the step after that must see the
      // rest update finished -> the panel stream also waits for ev_update before ITS trailing
      // columns are touched again.  That ordering is enforced at the top of the next iteration:
      // the update of step ob+1 is issued on sm, after ev_panel of that step.
    }
  }
  if (w.lookahead) {
    VC_CUDA_TRY(cudaEventRecord(w.ev_panel, sp));
    VC_CUDA_TRY(cudaStreamWaitEvent(sm, w.ev_panel, 0));
  }
}

// ---------------------------------------------------------------------------------------------
// roofline probes
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_dmma_probe(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
__global__ void k_fill(double2* p, int64_t n2, double v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n2; i += stride) p[i] = make_double2(v, v);
}
}  // namespace

extern "C" int vc_probe_dmma_peak(int32_t device, int32_t iters, double* tflops) {
  if (!tflops || iters <= 0) return VC_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return VC_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VC_ERR_CUDA;
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return VC_ERR_OOM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 2;
  k_dmma_probe<<<blocks, 256>>>(d, 64);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dmma_probe<<<blocks, 256>>>(d, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return VC_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = (double)blocks * 8.0 /*warps*/ * iters * 16.0 * 512.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return VC_OK;
}

extern "C" int vc_probe_hbm_write(int32_t device, int64_t bytes, int32_t iters, double* gbs) {
  if (!gbs || bytes < 16 || iters <= 0) return VC_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return VC_ERR_CUDA;
  double2* d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return VC_ERR_OOM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int64_t n2 = bytes / 16;
  k_fill<<<148 * 16, 256>>>(d, n2, 0.0);
  double best = 0.0;
  for (int rep = 0; rep < iters; ++rep) {
    cudaEventRecord(e0);
    k_fill<<<148 * 16, 256>>>(d, n2, 1.0 + rep);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return VC_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)(n2 * 16) / (ms * 1e-3) / 1e9;
    if (g > best) best = g;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *gbs = best;
  return VC_OK;
}
