// vc_chol.cu -- K2: bordered, blocked, right-looking fp64 Cholesky on sm_100a.
//
// Replaces base::chol (LAPACK dpotrf) at reference R-pkg/R/objective_functions.R:26 and, through
// the border rows, the triangular solves solve(t(R), X), solve(t(R), y), solve(t(R), res) of
// R-pkg/R/utils.R:94,100 and R-pkg/R/objective_functions.R:49.
//
// Storage: lower triangle of Sigma in a column-major lda x n_pad array `a`; right-hand sides live
// as ROWS of a separate column-major border array `b` (B^T, 128 rows per border tile).  Sweeping
// the factorisation over [a; b] turns the border into (L^-1 B)^T for free: every panel step
// treats a border tile as one more row tile (TRSM + update), it is never factored itself.
//
// Kernels:
//   k_potrf128     one CTA: factor a 128x128 diagonal tile in shared memory, and invert the factor.
//   k_trsm_panel   X = A * Linv^T for the row tiles below the diagonal tile: the panel solve as a
//                  128x128x128 DMMA GEMM (cp.async pipeline).
//   k_syrk_ws      trailing / inner update C -= P_I P_J^T.  Persistent (one CTA per SM walks a
//                  precomputed tile list), warp-specialised: one producer warp streams operand
//                  columns with cp.async.bulk (TMA engine, mbarrier complete_tx), eight consumer
//                  warps (64x32 warp tiles) issue DMMA m8n8k4 on the fp64 tensor pipe.
//   k_syrk_tiles   the same update, one CTA per tile with a cp.async pipeline (fallback / A-B).
#include "vc_common.cuh"
#include <limits.h>
#include <algorithm>
#include <stdlib.h>

namespace {

constexpr int TILE = VC_TILE;
constexpr int KC = 16;        // k-depth of one pipeline stage
constexpr int SLD = 132;      // shared-memory stride of one k-column: 128 + 4 doubles -> the
                              // (row, k) fragment pattern of m8n8k4 hits 16 distinct bank pairs
constexpr int STAGE_DOUBLES = 2 * KC * SLD;                       // A and B operand of one stage
constexpr int STAGE_BYTES = STAGE_DOUBLES * (int)sizeof(double);  // 33792
constexpr int GEMM_THREADS = 256;
constexpr int STAGES = 4;     // cp.async kernels
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES;                   // 135168 B
constexpr int WS_STAGES = 6;  // warp-specialised kernel
constexpr int WS_THREADS = 288;                                   // 8 consumer warps + 1 producer
constexpr int WS_SMEM = WS_STAGES * STAGE_BYTES + 2 * WS_STAGES * 8 + 16;
constexpr int SUPER = 16;     // super-block edge (tiles) of the L2-friendly tile order

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier / bulk-copy (TMA engine) helpers -------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col) on the fp64 tensor pipe (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

using GemmArgs = VcGemmArgs;

// local tile index of global tile t in a block-cyclic distribution over P processes
__device__ __forceinline__ int loc_tile(const GemmArgs& p, int t, int P) {
  return P > 1 ? ((t / p.blk) / P) * p.blk + t % p.blk : t;
}
// pointer to element (0, col) of row tile I in the (local) matrix / border; col in LOCAL elements
// (slot e of a batched launch: the matrix / border of evaluation e)
__device__ __forceinline__ double* row_tile_ptr(const GemmArgs& p, int e, int I, int64_t col, int64_t& ld) {
  if (I < p.nt) {
    double* t = vc_tile_ptr(p.a + e * p.a_es, p.cb, p.blk, loc_tile(p, I, p.pr), (int)(col >> 7), ld);
    return t + (col & (TILE - 1)) * ld;
  }
  ld = p.ldb;
  return p.b + e * p.b_es + (int64_t)(I - p.nt) * TILE + col * p.ldb;
}
// C tile (I, J), J a global column tile
__device__ __forceinline__ double* c_tile_ptr(const GemmArgs& p, int e, int I, int J, int64_t& ld) {
  return row_tile_ptr(p, e, I, (int64_t)loc_tile(p, J, p.pc) * TILE, ld);
}
// operand row tile I, panel column kcol: from a gathered panel buffer, or from the matrix itself
__device__ __forceinline__ const double* op_ptr(const GemmArgs& p, const VcOpSrc& s, int e, int I, int kcol,
                                                int64_t& ld) {
  if (s.pnl) {
    ld = s.ld;
    return s.pnl + (int64_t)vc_panel_pos(s.map, I, s.tile0) * s.tile_stride + (int64_t)(s.col0 + kcol) * s.ld;
  }
  return row_tile_ptr(p, e, I, (int64_t)p.k0 + kcol, ld);
}

// one pipeline stage of a warp: 4 k4 steps of 12 fragment loads + 32 DMMA.
// sa / sb already point at (warp row / col offset + g) of the stage's A / B operand.
__device__ __forceinline__ void warp_stage(const double* sa, const double* sb, int tig,
                                           double (&acc)[8][4][2]) {
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

// acc[i][j][0..1]: rows wm*64 + i*8 + (lane>>2), cols wn*32 + j*8 + 2*(lane&3) + {0,1}
__device__ __forceinline__ void gemm_mainloop(const double* __restrict__ Ag, int64_t lda_a,
                                              const double* __restrict__ Bg, int64_t lda_b, int K,
                                              double (&acc)[8][4][2], double* smem) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, tig = lane & 3;
  const int nk = K / KC;

  auto load_stage = [&](int slot, int kt) {
    double* sa = smem + slot * STAGE_DOUBLES;
    double* sb = sa + KC * SLD;
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
    const double* sa = smem + (kt % STAGES) * STAGE_DOUBLES;
    warp_stage(sa + wm * 64 + g, sa + KC * SLD + wn * 32 + g, tig, acc);
  }
  cp_async_wait<0>();
  __syncthreads();
}

// ---- panel solve: X(I, jt) = A(I, jt) * Linv^T ------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_trsm_panel(GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  const int I = p.tiles ? (p.tiles[blockIdx.x] & 0xffff)
                        : (((int)blockIdx.x < p.n_main) ? p.jt + 1 + p.i_skip + blockIdx.x : p.nt + (blockIdx.x - p.n_main));
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, tig = lane & 3;
  int64_t ld;
  const int e = blockIdx.y;
  double* Ct = c_tile_ptr(p, e, I, p.jt, ld);
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // the B operand element (n, k) is Linv[n + 128 k]
  gemm_mainloop(Ct, ld, p.linv + e * p.linv_es, TILE, TILE, acc, smem);
  double* cbase = Ct + (wm * 64 + g) + (int64_t)(wn * 32 + 2 * tig) * ld;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double* c = cbase + i * 8 + (int64_t)(j * 8) * ld;
      c[0] = acc[i][j][0];
      c[ld] = acc[i][j][1];
    }
}

// ---- update, one CTA per tile (fallback) ------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_syrk_tiles(GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  const int packed = p.tiles[blockIdx.x];
  const int I = packed & 0xffff, J = packed >> 16;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, tig = lane & 3;
  int64_t ldc, ldpi, ldpj;
  const int e = blockIdx.y;
  double* Ct = c_tile_ptr(p, e, I, J, ldc);
  const double* Pi = op_ptr(p, p.src_a, e, I, 0, ldpi);
  const double* Pj = op_ptr(p, p.src_b, e, J, 0, ldpj);
  double* cbase = Ct + (wm * 64 + g) + (int64_t)(wn * 32 + 2 * tig) * ldc;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  gemm_mainloop(Pi, ldpi, Pj, ldpj, p.K, acc, smem);
  // same arithmetic as k_syrk_ws (C + (-acc), one rounding) so the two kernels agree bit for bit
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double* c = cbase + i * 8 + (int64_t)(j * 8) * ldc;
      c[0] = c[0] + (-acc[i][j][0]);
      c[ldc] = c[ldc] + (-acc[i][j][1]);
    }
}

// ---- 32-row strips of a tile: remainder waves of the persistent update ------------------------
// A whole 128x128xK product occupies one SM for its full duration whatever the kernel; cut into
// four 32-row strips on four SMs it takes a quarter.  Same per-element DMMA order as the full-tile
// kernels, so results are bit-identical.  Block b: tile b / 4, strip b % 4.
constexpr int STRIP = 32;
__device__ __forceinline__ void warp_stage_strip(const double* sa, const double* sb, int tig, double (&acc)[4][2][2]) {
#pragma unroll
  for (int k4 = 0; k4 < KC / 4; ++k4) {
    const int kk = (k4 * 4 + tig) * SLD;
    double a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = sa[kk + i * 8];
#pragma unroll
    for (int j = 0; j < 2; ++j) b[j] = sb[kk + j * 8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}
// acc[i][j][0..1]: rows i*8 + (lane>>2) of the strip, cols warp*16 + j*8 + 2*(lane&3) + {0,1}
__device__ __forceinline__ void strip_mainloop(const double* __restrict__ Ag, int64_t lda_a,
                                               const double* __restrict__ Bg, int64_t lda_b, int K,
                                               double (&acc)[4][2][2], double* smem) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int nk = K / KC;
  auto load_stage = [&](int slot, int kt) {
    double* sa = smem + slot * STAGE_DOUBLES;
    double* sb = sa + KC * SLD;
    {
      const int col = tid >> 4, rc = (tid & 15) * 2;          // 16 k-columns x 32 rows of the strip
      cp_async16(sa + col * SLD + rc, Ag + (int64_t)(kt * KC + col) * lda_a + rc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = tid + GEMM_THREADS * i;
      const int col = id >> 6, rc = (id & 63) * 2;
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
    const double* sa = smem + (kt % STAGES) * STAGE_DOUBLES;
    warp_stage_strip(sa + g, sa + KC * SLD + warp * 16 + g, tig, acc);
  }
  cp_async_wait<0>();
  __syncthreads();
}
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_strip(GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  const int t = blockIdx.x >> 2, s = blockIdx.x & 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // C(I, J) -= P_I P_J^T on rows [32 s, 32 s + 32) of the tile; linear tile index lin_first + t over (slot, tile)
  const int lin = p.lin_first + t;
  const int e = lin / p.ntiles;
  const int packed = p.tiles[lin - e * p.ntiles];
  const int I = packed & 0xffff, J = packed >> 16;
  int64_t ldc, ldpi, ldpj;
  double* Ct = c_tile_ptr(p, e, I, J, ldc) + STRIP * s;
  const double* Pi = op_ptr(p, p.src_a, e, I, 0, ldpi) + STRIP * s;
  const double* Pj = op_ptr(p, p.src_b, e, J, 0, ldpj);
  strip_mainloop(Pi, ldpi, Pj, ldpj, p.K, acc, smem);
  double* cbase = Ct + g + (int64_t)(warp * 16 + 2 * tig) * ldc;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double* c = cbase + i * 8 + (int64_t)(j * 8) * ldc;
      red_add_f64(c, -acc[i][j][0]);
      red_add_f64(c + ldc, -acc[i][j][1]);
    }
}

// ---- update, persistent + warp-specialised ----------------------------------------------------
// shared: WS_STAGES x [A: KC x SLD | B: KC x SLD] doubles, then full[WS_STAGES], empty[WS_STAGES].
// Producer (warp 8): per stage, lane l streams k-column (l & 15) of operand (l >> 4) -- one
// 1-KB cp.async.bulk each -- into the slot, completion counted on full[slot].  Consumers wait on
// full[slot], run 4 k4 steps, and one lane per warp arrives on empty[slot].  No CTA-wide barrier
// in steady state: warps drift apart and keep the DMMA pipe fed across stage and tile boundaries.
__global__ void __launch_bounds__(WS_THREADS, 1) k_syrk_ws(GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WS_STAGES * STAGE_DOUBLES);
  uint64_t* empty = full + WS_STAGES;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < WS_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    // the TMA engine updates the barriers (complete_tx) through the async proxy: make the
    // initialisation visible there before the first bulk copy can land
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  const int nk = p.K / KC;
  // linear index over (evaluation slot, tile of the list); lin_stop > 0 cuts the walk short (the launcher
  // hands the last few indices to the strip kernel)
  const int total = p.lin_stop > 0 ? p.lin_stop : p.ntiles * p.nbatch;

  if (warp == 8) {
    // ===== producer =====
    unsigned it = 0;
    const int col = lane & 15, opnd = lane >> 4;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int e = t / p.ntiles;
      const int packed = p.tiles[t - e * p.ntiles];
      const int I = packed & 0xffff, J = packed >> 16;
      int64_t ldsrc;
      const double* src = opnd == 0 ? op_ptr(p, p.src_a, e, I, col, ldsrc) : op_ptr(p, p.src_b, e, J, col, ldsrc);
      // pull the NEXT tile's C into L2 while this tile computes (its loads then hit L2)
      const int tn = t + gridDim.x;
      if (tn < total) {
        const int en = tn / p.ntiles;
        const int pk = p.tiles[tn - en * p.ntiles];
        int64_t ldn;
        const double* cn = c_tile_ptr(p, en, pk & 0xffff, pk >> 16, ldn);
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_prefetch_l2(cn + (int64_t)(lane * 4 + c) * ldn, TILE * 8);
      }
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const unsigned slot = it % WS_STAGES;
        const unsigned round = it / WS_STAGES;
        mbar_wait(empty + slot, (round & 1) ^ 1);
        __syncwarp();
        if (lane == 0) mbar_expect_tx(full + slot, 2 * KC * TILE * 8);
        __syncwarp();
        double* dst = smem + slot * STAGE_DOUBLES + opnd * (KC * SLD) + col * SLD;
        bulk_g2s(dst, src + (int64_t)kt * KC * ldsrc, TILE * 8, full + slot);
      }
    }
  } else {
    // ===== consumers =====
    const int wm = warp & 1, wn = warp >> 1;
    const int g = lane >> 2, tig = lane & 3;
    unsigned it = 0;
    int packed_next = ((int)blockIdx.x < total) ? p.tiles[blockIdx.x % p.ntiles] : 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int packed = packed_next;
      const int e = t / p.ntiles;
      // the next tile's index is fetched a whole tile ahead: its ~1 us global-load latency would
      // otherwise sit on the critical path of every tile boundary
      if (t + (int)gridDim.x < total) packed_next = p.tiles[(t + gridDim.x) % p.ntiles];
      const int I = packed & 0xffff, J = packed >> 16;
      int64_t ldc;
      double* Ct = c_tile_ptr(p, e, I, J, ldc);
      double* cbase = Ct + (wm * 64 + g) + (int64_t)(wn * 32 + 2 * tig) * ldc;
      double acc[8][4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const unsigned slot = it % WS_STAGES;
        mbar_wait(full + slot, (it / WS_STAGES) & 1);
        __syncwarp();   // lanes may leave the polling loop in different iterations; mma.sync needs convergence
        const double* sa = smem + slot * STAGE_DOUBLES;
        warp_stage(sa + wm * 64 + g, sa + KC * SLD + wn * 32 + g, tig, acc);
        // Order this warp's generic-proxy reads of the slot (the LDS above) before the
        // async-proxy refill the arrive allows.  Without the fence ptxas hoists the arrive right
        // behind the ISSUE of the last LDS and the TMA engine can overwrite data still being read
        // (observed: nondeterministic ~1e-5 errors; see tests: ..._bitwise_equals_per_tile_kernel).
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
      }
      // epilogue: C -= acc as fire-and-forget fp64 reductions resolved in L2 (RED.ADD.F64).  No
      // load of C on the SM: the tile boundary costs no round trip to memory, and every element
      // receives exactly one reduction per launch, so the result stays deterministic.
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double* c = cbase + i * 8 + (int64_t)(j * 8) * ldc;
          red_add_f64(c, -acc[i][j][0]);
          red_add_f64(c + ldc, -acc[i][j][1]);
        }
    }
  }
}

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
k_potrf128(double* At, int64_t lda, int jt, double* linv_all, double* logdet_part, int* info, VcPotrfBatch bat) {
  At += blockIdx.x * bat.a_es; linv_all += blockIdx.x * bat.linv_es; logdet_part += blockIdx.x * bat.nt; info += blockIdx.x;
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


// ---------------------------------------------------------------------------------------------
// k_potrf128b: blocked version of the diagonal-tile factorisation (default).  The 128x128 tile is
// treated as 8x8 blocks of 16x16.  Per block column: warp 0 factors the 16x16 diagonal block and
// inverts it in registers (shuffles); all warps solve the 16-wide panel below (DMMA, multiply by
// the block inverse) and apply the rank-16 update to the trailing lower triangle (DMMA on 8x8
// sub-tiles).  The full inverse is then assembled block sub-diagonal by block sub-diagonal:
//   X_ij = -X_ii * sum_{k=j}^{i-1} L_ik X_kj      (all 16x16x16 DMMA products, one warp per block)
// Shared array M: 129 columns x 132 doubles (column stride 132: conflict-free m8n8k4 fragments)
//   L(i, j) (i >= j) at M[j*PLD + i];   X(i, c) = Linv(i, c) (i >= c) at M[(i+1)*PLD + c].
// ---------------------------------------------------------------------------------------------
constexpr int PLD = 132;
constexpr int PB = 16;
constexpr int SS_LD = 20;   // per-warp 16x16 scratch, column stride 20 doubles
constexpr int POTRFB_SMEM = ((TILE + 1) * PLD + TILE + 8 * PB * SS_LD + 2 * PB * 17) * (int)sizeof(double);

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <bool FACTOR>   // FACTOR = false: the tile already holds L; only invert it (vc_gls_chol)
__global__ void __launch_bounds__(POTRF_THREADS, 1)
k_potrf128b(double* At, int64_t lda, int jt, double* linv_all, double* logdet_part, int* info, VcPotrfBatch bat) {
  At += blockIdx.x * bat.a_es; linv_all += blockIdx.x * bat.linv_es; logdet_part += blockIdx.x * bat.nt; info += blockIdx.x;
  extern __shared__ __align__(16) double M[];
  double* sdiag = M + (TILE + 1) * PLD;
  double* scratch = sdiag + TILE;
  double* Dsh = scratch + 8 * PB * SS_LD;   // 16 x 17
  double* Xsh = Dsh + PB * 17;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;

  for (int idx = tid; idx < TILE * TILE; idx += POTRF_THREADS) {
    const int r = idx & (TILE - 1), c = idx >> 7;
    if (r >= c) M[c * PLD + r] = At[r + (int64_t)c * lda];
  }
  __syncthreads();

  for (int kb = 0; kb < TILE / PB; ++kb) {
    const int k0 = kb * PB;
    // ---- A: diagonal 16x16 block, factor + inverse; thread (i, j) owns element (i, j) ----
    {
      const int i = tid & 15, j = tid >> 4;
      Dsh[i * 17 + j] = (i >= j) ? M[(k0 + j) * PLD + k0 + i] : 0.0;
      Xsh[i * 17 + j] = (i == j) ? 1.0 : 0.0;
      __syncthreads();
      for (int jj = 0; jj < PB; ++jj) {
        const double ajj = Dsh[jj * 17 + jj];              // final: columns < jj were applied before the barrier
        if (FACTOR) {
          // the pivot element itself is left untouched inside the loop (sqrt goes to sdiag), so one
          // barrier per phase is enough: scale column jj | barrier | rank-1 update | barrier
          if (j == jj && i > jj) Dsh[i * 17 + jj] *= 1.0 / sqrt(ajj);
          if (tid == jj * 17) {                              // thread (jj, jj)
            sdiag[k0 + jj] = sqrt(ajj);
            if (!(ajj > 0.0)) atomicMin(info, jt * TILE + k0 + jj + 1);
          }
          __syncthreads();
          if (j > jj && i >= j) Dsh[i * 17 + j] = fma(-Dsh[i * 17 + jj], Dsh[j * 17 + jj], Dsh[i * 17 + j]);
          __syncthreads();
        } else if (tid == jj * 17) {
          sdiag[k0 + jj] = ajj;
          if (!(ajj > 0.0)) atomicMin(info, jt * TILE + k0 + jj + 1);
        }
      }
      __syncthreads();
      if (FACTOR && i == j) Dsh[i * 17 + i] = sdiag[k0 + i];
      __syncthreads();
      // X = D^-1, right-looking forward substitution on the identity (element (i, c = j))
      for (int k = 0; k < PB; ++k) {
        if (i == k) Xsh[k * 17 + j] *= 1.0 / Dsh[k * 17 + k];
        __syncthreads();
        if (i > k) Xsh[i * 17 + j] = fma(-Dsh[i * 17 + k], Xsh[k * 17 + j], Xsh[i * 17 + j]);
        __syncthreads();
      }
      if (i >= j) {
        if (FACTOR) M[(k0 + j) * PLD + k0 + i] = Dsh[i * 17 + j];
        M[(k0 + i + 1) * PLD + k0 + j] = Xsh[i * 17 + j];
      }
    }
    __syncthreads();
    const int nsub = (TILE - k0 - PB) / 8;   // 8-row sub-tiles below the diagonal block
    // ---- B: panel solve  P = A * Dinv^T  on rows [k0+16, 128) ----
    for (int ms = warp; FACTOR && ms < nsub; ms += 8) {
      const int row0 = k0 + PB + 8 * ms;
      double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int k = 4 * k4 + tig;
        const double af = M[(k0 + k) * PLD + row0 + g];
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          const int c = 8 * nn + g;
          const double bf = (k <= c) ? M[(k0 + c + 1) * PLD + k0 + k] : 0.0;
          dmma884(acc[nn][0], acc[nn][1], af, bf);
        }
      }
      __syncwarp();
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        M[(k0 + 8 * nn + 2 * tig) * PLD + row0 + g] = acc[nn][0];
        M[(k0 + 8 * nn + 2 * tig + 1) * PLD + row0 + g] = acc[nn][1];
      }
    }
    __syncthreads();
    // ---- C: trailing update on the lower triangle of 8x8 sub-tiles ----
    const int npair = nsub * (nsub + 1) / 2;
    for (int t = warp; FACTOR && t < npair; t += 8) {
      int sa = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
      while ((sa + 1) * (sa + 2) / 2 <= t) ++sa;
      while (sa * (sa + 1) / 2 > t) --sa;
      const int sb = t - sa * (sa + 1) / 2;
      const int rowA = k0 + PB + 8 * sa, rowB = k0 + PB + 8 * sb;
      double c0 = M[(rowB + 2 * tig) * PLD + rowA + g];
      double c1 = M[(rowB + 2 * tig + 1) * PLD + rowA + g];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int k = k0 + 4 * k4 + tig;
        dmma884(c0, c1, -M[k * PLD + rowA + g], M[k * PLD + rowB + g]);
      }
      if (sa != sb || g >= 2 * tig) M[(rowB + 2 * tig) * PLD + rowA + g] = c0;
      if (sa != sb || g >= 2 * tig + 1) M[(rowB + 2 * tig + 1) * PLD + rowA + g] = c1;
    }
    __syncthreads();
  }

  // ---- inverse, block sub-diagonal d = 1..7: X_ij = -X_ii * sum_k L_ik X_kj ----
  double* ss = scratch + warp * PB * SS_LD;
  for (int d = 1; d < TILE / PB; ++d) {
    const int bi = d + warp, bj = warp;
    if (bi < TILE / PB) {
      double s[2][2][2];
#pragma unroll
      for (int mm = 0; mm < 2; ++mm)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) s[mm][nn][0] = s[mm][nn][1] = 0.0;
      for (int kbk = bj; kbk < bi; ++kbk) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int kk = 4 * k4 + tig;             // within block kbk
          double af[2], bf[2];
#pragma unroll
          for (int mm = 0; mm < 2; ++mm) af[mm] = M[(PB * kbk + kk) * PLD + PB * bi + 8 * mm + g];
#pragma unroll
          for (int nn = 0; nn < 2; ++nn) {
            const int n = 8 * nn + g;
            // X_kj(kk, n); the diagonal block (kbk == bj) is lower triangular, its upper part is not stored
            bf[nn] = (kbk > bj || kk >= n) ? M[(PB * kbk + kk + 1) * PLD + PB * bj + n] : 0.0;
          }
#pragma unroll
          for (int mm = 0; mm < 2; ++mm)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) dmma884(s[mm][nn][0], s[mm][nn][1], af[mm], bf[nn]);
        }
      }
      // S -> per-warp scratch as S(kk, n) at ss[n * SS_LD + kk]
#pragma unroll
      for (int mm = 0; mm < 2; ++mm)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          ss[(8 * nn + 2 * tig) * SS_LD + 8 * mm + g] = s[mm][nn][0];
          ss[(8 * nn + 2 * tig + 1) * SS_LD + 8 * mm + g] = s[mm][nn][1];
        }
      __syncwarp();
      double xo[2][2][2];
#pragma unroll
      for (int mm = 0; mm < 2; ++mm)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) xo[mm][nn][0] = xo[mm][nn][1] = 0.0;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int kk = 4 * k4 + tig;
        double af[2], bf[2];
#pragma unroll
        for (int mm = 0; mm < 2; ++mm) {
          const int r = 8 * mm + g;                // -X_ii(r, kk), lower triangular
          af[mm] = (kk <= r) ? -M[(PB * bi + r + 1) * PLD + PB * bi + kk] : 0.0;
        }
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) bf[nn] = ss[(8 * nn + g) * SS_LD + kk];
#pragma unroll
        for (int mm = 0; mm < 2; ++mm)
#pragma unroll
          for (int nn = 0; nn < 2; ++nn) dmma884(xo[mm][nn][0], xo[mm][nn][1], af[mm], bf[nn]);
      }
      // X_ij(r, c) -> M[(16 bi + r + 1) * PLD + 16 bj + c]
#pragma unroll
      for (int mm = 0; mm < 2; ++mm)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          double* dst = M + (PB * bi + 8 * mm + g + 1) * PLD + PB * bj + 8 * nn + 2 * tig;
          dst[0] = xo[mm][nn][0];
          dst[1] = xo[mm][nn][1];
        }
    }
    __syncthreads();
  }

  // ---- write back ----
  double* linv = linv_all + (int64_t)jt * TILE * TILE;
  for (int idx = tid; idx < TILE * TILE; idx += POTRF_THREADS) {
    const int r = idx & (TILE - 1), c = idx >> 7;
    if (FACTOR && r >= c) At[r + (int64_t)c * lda] = M[c * PLD + r];
    linv[idx] = (r >= c) ? M[(r + 1) * PLD + c] : 0.0;
  }
  if (tid < 32) {
    double sl = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) sl += log(sdiag[lane * 4 + a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
    if (lane == 0) logdet_part[jt] = sl;
  }
}

// ---------------------------------------------------------------------------------------------
// k_potrf128c (default): latency-oriented version of k_potrf128b, same storage and same outputs.
// The chain of 128 pivots bounds this kernel, so
//  * the 16x16 diagonal block is factored AND inverted by one warp entirely in registers: lane i
//    owns row i, pivots and column entries travel by shuffles, 1/sqrt by rsqrt.approx + Newton --
//    no shared-memory round trip and no CTA barrier inside the 16 pivot steps (~135 cycles/pivot);
//  * while warp 0 does that, warps 1..7 finish the previous rank-16 update and assemble the block
//    row of the inverse that became computable (X_ij = -X_ii sum_k L_ik X_kj), so neither sits on
//    the critical path; only the three 8x8 pairs of the next diagonal block are updated first;
//  * tile load / store are 16-byte, the inverse is written through 4x8 patches (full 32-byte
//    sectors in global memory, conflict-free in shared memory).
// ---------------------------------------------------------------------------------------------
constexpr int POTRFC_SMEM = ((TILE + 1) * PLD + TILE + 8 * PB * SS_LD) * (int)sizeof(double);
#ifdef VC_POTRF_CLOCKS   // tools/potrf_probe.cu: cycle stamps at the phase boundaries
__device__ long long g_potrf_clk[64];
#define POTRF_STAMP(k) do { if (threadIdx.x == 0) g_potrf_clk[k] = clock64(); __syncwarp(); } while (0)
#else
#define POTRF_STAMP(k) do { } while (0)
#endif

// rank-16 update of the 8x8 pair t of the trailing lower triangle behind block column k0 / 16
__device__ __forceinline__ void potrf_update_pair(double* M, int k0, int t, int g, int tig) {
  int sa = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
  while ((sa + 1) * (sa + 2) / 2 <= t) ++sa;
  while (sa * (sa + 1) / 2 > t) --sa;
  const int sb = t - sa * (sa + 1) / 2;
  const int rowA = k0 + PB + 8 * sa, rowB = k0 + PB + 8 * sb;
  double c0 = M[(rowB + 2 * tig) * PLD + rowA + g];
  double c1 = M[(rowB + 2 * tig + 1) * PLD + rowA + g];
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    const int k = k0 + 4 * k4 + tig;
    dmma884(c0, c1, -M[k * PLD + rowA + g], M[k * PLD + rowB + g]);
  }
  if (sa != sb || g >= 2 * tig) M[(rowB + 2 * tig) * PLD + rowA + g] = c0;
  if (sa != sb || g >= 2 * tig + 1) M[(rowB + 2 * tig + 1) * PLD + rowA + g] = c1;
}

// one 16x16 block of the inverse: X_ij = -X_ii * sum_{k=j}^{i-1} L_ik X_kj (one warp, DMMA)
__device__ __forceinline__ void potrf_xrow_item(double* M, double* ss, int bi, int bj, int g, int tig) {
  double s[2][2][2];
#pragma unroll
  for (int mm = 0; mm < 2; ++mm)
#pragma unroll
    for (int nn = 0; nn < 2; ++nn) s[mm][nn][0] = s[mm][nn][1] = 0.0;
  for (int kbk = bj; kbk < bi; ++kbk) {
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const int kk = 4 * k4 + tig;
      double af[2], bf[2];
#pragma unroll
      for (int mm = 0; mm < 2; ++mm) af[mm] = M[(PB * kbk + kk) * PLD + PB * bi + 8 * mm + g];
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        const int n = 8 * nn + g;
        bf[nn] = (kbk > bj || kk >= n) ? M[(PB * kbk + kk + 1) * PLD + PB * bj + n] : 0.0;
      }
#pragma unroll
      for (int mm = 0; mm < 2; ++mm)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) dmma884(s[mm][nn][0], s[mm][nn][1], af[mm], bf[nn]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int mm = 0; mm < 2; ++mm)
#pragma unroll
    for (int nn = 0; nn < 2; ++nn) {
      ss[(8 * nn + 2 * tig) * SS_LD + 8 * mm + g] = s[mm][nn][0];
      ss[(8 * nn + 2 * tig + 1) * SS_LD + 8 * mm + g] = s[mm][nn][1];
    }
  __syncwarp();
  double xo[2][2][2];
#pragma unroll
  for (int mm = 0; mm < 2; ++mm)
#pragma unroll
    for (int nn = 0; nn < 2; ++nn) xo[mm][nn][0] = xo[mm][nn][1] = 0.0;
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    const int kk = 4 * k4 + tig;
    double af[2], bf[2];
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
      const int r = 8 * mm + g;
      af[mm] = (kk <= r) ? -M[(PB * bi + r + 1) * PLD + PB * bi + kk] : 0.0;
    }
#pragma unroll
    for (int nn = 0; nn < 2; ++nn) bf[nn] = ss[(8 * nn + g) * SS_LD + kk];
#pragma unroll
    for (int mm = 0; mm < 2; ++mm)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) dmma884(xo[mm][nn][0], xo[mm][nn][1], af[mm], bf[nn]);
  }
#pragma unroll
  for (int mm = 0; mm < 2; ++mm)
#pragma unroll
    for (int nn = 0; nn < 2; ++nn) {
      double* dst = M + (PB * bi + 8 * mm + g + 1) * PLD + PB * bj + 8 * nn + 2 * tig;
      dst[0] = xo[mm][nn][0];
      dst[1] = xo[mm][nn][1];
    }
}

// 16x16 diagonal block at (k0, k0): factor (FACTOR) and invert, one warp.  Lane i (and its mirror
// i + 16) owns row i in registers.  The dependent chain per pivot is
//   1/sqrt (MUFU + 3rd-order correction) -> scale own column entry -> own diagonal -> shuffle,
// everything else (the rank-1 update of the other columns) reads the scaled column back from
// shared memory (its final location in M) as broadcast 16-byte loads and fills the issue slots
// the chain leaves free.  Template recursion keeps every register-array index a compile-time
// constant (a partially unrolled loop would push the arrays to local memory).
struct DiagRegs {
  double a[PB], invd[PB], x[PB];
  double dg, mydiag;
  int bad;
};
// 1 / sqrt(s): rsqrt.approx (~2^-22) + one third-order step r (1 + e/2 + 3 e^2 / 8), e = 1 - s r^2
__device__ __forceinline__ double rsqrt3(double s) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s));
  const double t = s * y0;
  const double e = fma(-t, y0, 1.0);
  const double c = fma(0.375, e, 0.5);
  return fma(y0 * e, c, y0);
}
// Column values fetched in step J are consumed in step J + 1 (software pipelining by one step): the
// loads have a whole pivot chain of time to land, so the single warp never stalls on them.
struct DiagCol { double v[PB]; };
template <int J, int K>
__device__ __forceinline__ void potrf_diag_fetch(DiagCol& C, const double* Mb) {   // C.v[K] = L(K, J), K >= J + 2
  if constexpr (K < PB) {
    if constexpr (K % 2 == 0 && K + 1 < PB) {
      const double2 v = *reinterpret_cast<const double2*>(Mb + J * PLD + K);
      C.v[K] = v.x;
      C.v[K + 1] = v.y;
      potrf_diag_fetch<J, K + 2>(C, Mb);
    } else {
      C.v[K] = Mb[J * PLD + K];
      potrf_diag_fetch<J, K + 1>(C, Mb);
    }
  }
}
template <int K>
__device__ __forceinline__ void potrf_diag_apply(DiagRegs& R, const DiagCol& C, double lprev) {
  if constexpr (K < PB) {
    R.a[K] = fma(-lprev, C.v[K], R.a[K]);
    potrf_diag_apply<K + 1>(R, C, lprev);
  }
}
// Pivot J from its value p: r = 1 / sqrt(p) for everyone, sqrt(p) kept by lane J
__device__ __forceinline__ double potrf_diag_pivot(DiagRegs& R, int J, int i, double p) {
  if (!(p > 0.0)) R.bad = min(R.bad, J);
  const double r = rsqrt3(p);
  double sq = p * r;
  sq = fma(fma(-sq, sq, p), 0.5 * r, sq);
  R.mydiag = (i == J) ? sq : R.mydiag;
  return r;
}
// One pivot step.  On entry r = 1 / L(J, J); u = A'(J+1, J) and q = A'(J+1, J+1) are the not yet
// scaled column-J entry and the diagonal of row J+1 (updated through column J-1), broadcast one
// step ahead, so that EVERY lane can form L(J+1, J) = u r and the next pivot q - L(J+1, J)^2 itself:
// the dependent chain from one pivot to the next is  mul -> fma -> rsqrt  with no shuffle in it.
// Cp = column J-1 below row J (fetched in the previous step), lprev = L(i, J-1).
// R.a[J] arrives masked (0 for lanes i <= J).
template <int J>
__device__ __forceinline__ void potrf_diag_step(DiagRegs& R, double* Mb, int i, bool st, double r, double u,
                                                double q, const DiagCol& Cp, double lprev) {
  if constexpr (J < PB) {
    R.invd[J] = r;
    const double lij = R.a[J] * r;
    R.dg = fma(-lij, lij, R.dg);
    double rn = 0.0, un = 0.0, qn = 0.0;
    if constexpr (J + 1 < PB) {
      const double l1 = u * r;                              // L(J+1, J), the arithmetic lane J+1 does itself
      rn = potrf_diag_pivot(R, J + 1, i, fma(-l1, l1, q));
      if constexpr (J >= 1) R.a[J + 1] = fma(-lprev, Cp.v[J + 1], R.a[J + 1]);   // pending column J-1
      const double t = fma(-lij, l1, R.a[J + 1]);
      R.a[J + 1] = (i > J + 1) ? t : 0.0;
      if constexpr (J + 2 < PB) {
        un = shfl_d(R.a[J + 1], J + 2);
        qn = shfl_d(R.dg, J + 2);
      }
    }
    if (st && i > J) Mb[J * PLD + i] = lij;
    if constexpr (J >= 1) potrf_diag_apply<J + 2>(R, Cp, lprev);   // the rest of pending column J-1
    __syncwarp();
    DiagCol Cn;
    potrf_diag_fetch<J, J + 2>(Cn, Mb);
    potrf_diag_step<J + 1>(R, Mb, i, st, rn, un, qn, Cn, lij);
  }
}
// X = D^-1: lane c carries column c (x = e_c) through a right-looking forward substitution.  Column
// K of D (contiguous in shared memory) is fetched one step ahead of its use, so the loads never
// stall the single warp; the dependent chain is mul -> fma per column.
template <int K, int I>
__device__ __forceinline__ void potrf_diag_inv_apply(DiagRegs& R, const DiagCol& C) {
  if constexpr (I < PB) {
    R.x[I] = fma(-C.v[I], R.x[K], R.x[I]);
    potrf_diag_inv_apply<K, I + 1>(R, C);
  }
}
template <int K>
__device__ __forceinline__ void potrf_diag_inv(DiagRegs& R, const double* Mb, const DiagCol& Cc) {
  if constexpr (K < PB) {
    DiagCol Cn;
    if constexpr (K + 1 < PB) potrf_diag_fetch<K + 1, K + 2>(Cn, Mb);   // column K+1 below its diagonal
    R.x[K] *= R.invd[K];
    potrf_diag_inv_apply<K, K + 1>(R, Cc);
    potrf_diag_inv<K + 1>(R, Mb, Cn);
  }
}
template <int I>
__device__ __forceinline__ void potrf_diag_inv_init(DiagRegs& R, int c) {
  if constexpr (I < PB) {
    R.x[I] = (I == c) ? 1.0 : 0.0;
    potrf_diag_inv_init<I + 1>(R, c);
  }
}
// X(J, c = i) -> M, entries above the diagonal go to a dump slot: an address select instead of a
// predicate, because the compiler turns sixteen `if (J >= i)` stores into a divergent branch cascade
template <int J>
__device__ __forceinline__ void potrf_diag_store_x(const DiagRegs& R, double* Mb, double* dump, int i) {
  if constexpr (J < PB) {
    double* dst = (J >= i) ? Mb + (J + 1) * PLD + i : dump;
    *dst = R.x[J];
    potrf_diag_store_x<J + 1>(R, Mb, dump, i);
  }
}
template <int J>
__device__ __forceinline__ void potrf_diag_load(DiagRegs& R, const double* Mb, int i) {
  if constexpr (J < PB) {
    const double v = Mb[J * PLD + max(i, J)];                 // in-bounds for every lane
    R.a[J] = (J < i) ? v : 0.0;
    potrf_diag_load<J + 1>(R, Mb, i);
  }
}
template <int J>
__device__ __forceinline__ void potrf_diag_recips(DiagRegs& R) {   // invert-only path
  if constexpr (J < PB) {
    const double p = shfl_d(R.dg, J);
    if (!(p > 0.0)) R.bad = min(R.bad, J);
    R.invd[J] = 1.0 / p;
    potrf_diag_recips<J + 1>(R);
  }
}
template <bool FACTOR>
__device__ __forceinline__ int potrf_diag_block(double* M, double* sdiag, double* dump, int k0, int lane) {
  const int i = lane & 15;
  const bool st = lane < PB;
  double* Mb = M + k0 * PLD + k0;
  DiagRegs R;
  R.dg = Mb[i * PLD + i];
  R.mydiag = R.dg;
  R.bad = PB;
  if constexpr (FACTOR) {
    potrf_diag_load<0>(R, Mb, i);
    POTRF_STAMP(40);
    const double r0 = potrf_diag_pivot(R, 0, i, shfl_d(R.dg, 0));
    const double u0 = shfl_d(R.a[0], 1), q0 = shfl_d(R.dg, 1);
    DiagCol C0;
#ifndef VC_POTRF_SKIP_FACTOR
    potrf_diag_step<0>(R, Mb, i, st, r0, u0, q0, C0, 0.0);
#endif
    if (st) Mb[i * PLD + i] = R.mydiag;
  } else {
    potrf_diag_recips<0>(R);
  }
  __syncwarp();
  POTRF_STAMP(41);
#ifndef VC_POTRF_SKIP_INV
  {
    DiagCol C0;
    potrf_diag_fetch<0, 1>(C0, Mb);
    potrf_diag_inv_init<0>(R, i);
    potrf_diag_inv<0>(R, Mb, C0);
  }
#endif
  POTRF_STAMP(42);
  if (st) {
#ifndef VC_POTRF_SKIP_INV
    potrf_diag_store_x<0>(R, Mb, dump + lane, i);
#endif
    sdiag[k0 + i] = R.mydiag;
  }
  return R.bad;
}

// helper-warp stores (threads ht = 0 .. nh-1): one 16-column block of L, one 16-row block of the inverse
__device__ __forceinline__ void potrf_store_L_cols(const double* M, double* At, int64_t lda, int kb, int ht, int nh) {
  const int c0 = kb * PB, r0 = kb * PB;                     // rows r0 .. 127, columns c0 .. c0+15
  const int nr2 = (TILE - r0) / 2;
  for (int idx = ht; idx < nr2 * PB; idx += nh) {
    const int r = r0 + 2 * (idx % nr2), c = c0 + idx / nr2;
    if (r >= c) *reinterpret_cast<double2*>(At + r + (int64_t)c * lda) = *reinterpret_cast<const double2*>(M + c * PLD + r);
    else if (r + 1 == c) At[r + 1 + (int64_t)c * lda] = M[c * PLD + r + 1];
  }
}
__device__ __forceinline__ void potrf_store_X_rows(const double* M, double* linv, int bi, int hw, int nhw, int lane) {
  // 16 (rows) x 4 (columns) patches, 16 bytes per lane: four full 128-byte lines per store.  Only the
  // block columns <= bi are written (the buffer is zero-filled once when it is allocated); the upper
  // half of the diagonal block is written as zeros.
  const int r = PB * bi + 2 * (lane & 7);
  for (int pt = hw; pt < 4 * (bi + 1); pt += nhw) {
    const int c = 4 * pt + (lane >> 3);
    double2 v;
    v.x = (r >= c) ? M[(r + 1) * PLD + c] : 0.0;
    v.y = (r + 1 >= c) ? M[(r + 2) * PLD + c] : 0.0;
    *reinterpret_cast<double2*>(linv + r + TILE * c) = v;
  }
}

template <bool FACTOR>
__global__ void __launch_bounds__(POTRF_THREADS, 1)
k_potrf128c(double* At, int64_t lda, int jt, double* linv_all, double* logdet_part, int* info, VcPotrfBatch bat) {
  At += blockIdx.x * bat.a_es; linv_all += blockIdx.x * bat.linv_es; logdet_part += blockIdx.x * bat.nt; info += blockIdx.x;
  extern __shared__ __align__(16) double M[];
  double* sdiag = M + (TILE + 1) * PLD;
  double* scratch = sdiag + TILE;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  double* ss = scratch + warp * PB * SS_LD;
  double* linv = linv_all + (int64_t)jt * TILE * TILE;
  POTRF_STAMP(0);

  // load: 16-byte async copies, first block column as its own group so that the first diagonal
  // block can start before the rest of the tile has arrived.  Chunks that straddle the diagonal
  // bring one upper-triangle word along; it lands on a slot of the inverse that is written before
  // it is ever read.
  for (int idx = tid; idx < (TILE / 2) * PB; idx += POTRF_THREADS) {
    const int r = (idx & 63) * 2, c = idx >> 6;
    if (r + 1 >= c) cp_async16(M + c * PLD + r, At + r + (int64_t)c * lda);
  }
  cp_async_commit();
  for (int idx = tid + (TILE / 2) * PB; idx < TILE * TILE / 2; idx += POTRF_THREADS) {
    const int r = (idx & 63) * 2, c = idx >> 6;
    if (r + 1 >= c) cp_async16(M + c * PLD + r, At + r + (int64_t)c * lda);
  }
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  POTRF_STAMP(1);

  for (int kb = 0; kb < TILE / PB; ++kb) {
    const int k0 = kb * PB;
    // phase 1: the previous rank-16 update on the next diagonal block only (pairs 0, 1, 2)
    if (FACTOR && kb > 0) {
      if (warp < 3) potrf_update_pair(M, k0 - PB, warp, g, tig);
      __syncthreads();
    }
    POTRF_STAMP(2 + 4 * kb);
    // phase 2: warp 0 factors + inverts the diagonal block; the others finish update kb-1, assemble
    // block row kb-1 of the inverse and write finished parts of L and of the inverse to global memory
    if (warp == 0) {
      const int bad = potrf_diag_block<FACTOR>(M, sdiag, ss, k0, lane);
      if (bad < PB && lane == 0) atomicMin(info, jt * TILE + k0 + bad + 1);
      POTRF_STAMP(3 + 4 * kb);
    } else if (kb > 0) {
      const int nxr = kb - 1;
      const int nsubp = (TILE - k0) / 8;                    // 8-row sub-tiles behind block column kb-1
      const int nitems = nxr + (FACTOR ? nsubp * (nsubp + 1) / 2 - 3 : 0);
      // stores first: they drain while the DMMA items below run
      if (FACTOR) potrf_store_L_cols(M, At, lda, kb - 1, tid - 32, POTRF_THREADS - 32);
      if (kb > 1) potrf_store_X_rows(M, linv, kb - 2, warp - 1, 7, lane);
      for (int t = warp - 1; t < nitems; t += 7) {
        if (t < nxr) potrf_xrow_item(M, ss, kb - 1, t, g, tig);
        else potrf_update_pair(M, k0 - PB, t - nxr + 3, g, tig);
      }
    }
    if (kb == 0) cp_async_wait<0>();                        // the rest of the tile (every thread its own copies)
    __syncthreads();
    POTRF_STAMP(4 + 4 * kb);
    // phase 3: panel solve  P = A * Dinv^T  on rows [k0+16, 128)
    if (FACTOR) {
      const int nsub = (TILE - k0 - PB) / 8;
      for (int ms = warp; ms < nsub; ms += 8) {
        const int row0 = k0 + PB + 8 * ms;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int k = 4 * k4 + tig;
          const double af = M[(k0 + k) * PLD + row0 + g];
#pragma unroll
          for (int nn = 0; nn < 2; ++nn) {
            const int c = 8 * nn + g;
            const double bf = (k <= c) ? M[(k0 + c + 1) * PLD + k0 + k] : 0.0;
            dmma884(acc[nn][0], acc[nn][1], af, bf);
          }
        }
        __syncwarp();
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          M[(k0 + 8 * nn + 2 * tig) * PLD + row0 + g] = acc[nn][0];
          M[(k0 + 8 * nn + 2 * tig + 1) * PLD + row0 + g] = acc[nn][1];
        }
      }
      __syncthreads();
    }
    POTRF_STAMP(5 + 4 * kb);
  }
  // last block row of the inverse
  if (warp < TILE / PB - 1) potrf_xrow_item(M, ss, TILE / PB - 1, warp, g, tig);
  __syncthreads();
  POTRF_STAMP(34);

  // ---- what is left to write: the last block column of L, the last two block rows of the inverse ----
  if (FACTOR) potrf_store_L_cols(M, At, lda, TILE / PB - 1, tid, POTRF_THREADS);
  POTRF_STAMP(35);
  potrf_store_X_rows(M, linv, TILE / PB - 2, warp, POTRF_THREADS / 32, lane);
  potrf_store_X_rows(M, linv, TILE / PB - 1, warp, POTRF_THREADS / 32, lane);
  POTRF_STAMP(36);
  if (tid < 32) {
    double sl = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) sl += log(sdiag[lane * 4 + a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
    if (lane == 0) logdet_part[jt] = sl;
  }
}

// copy K panel columns (local column lcol0..) of the listed row tiles into a tile-major buffer:
// tile I -> dst + (I - tile0) * 128 * K, column-major inside with ld 128 (distributed driver)
__global__ void __launch_bounds__(256) k_pack_tiles(GemmArgs p, const int* __restrict__ tiles, int lcol0, int K,
                                                    double* __restrict__ dst, int tile0, VcPanelMap map) {
  const int I = tiles[blockIdx.x] & 0xffff;
  int64_t ld;
  const double* src = row_tile_ptr(p, 0, I, lcol0, ld);
  double* d = dst + (int64_t)vc_panel_pos(map, I, tile0) * TILE * K;
  for (int idx = threadIdx.x; idx < (TILE / 2) * K; idx += blockDim.x) {
    const int r2 = idx % (TILE / 2), c = idx / (TILE / 2);
    *reinterpret_cast<double2*>(d + (int64_t)c * TILE + 2 * r2) =
        *reinterpret_cast<const double2*>(src + (int64_t)c * ld + 2 * r2);
  }
}

__global__ void k_reset_info(int* info) { info[blockIdx.x] = INT_MAX; }

// ---- host: tile lists -------------------------------------------------------------------------
// lower trapezoid J in [J0, J1), I in [J, nrows): valid tiles only, super-block by super-block
// (row of super-blocks by row), I fastest inside a super-block, so that the ~148 tiles in flight
// share a handful of panel row blocks in L2.
void append_trapezoid(std::vector<int>& out, int J0, int J1, int nrows) {
  const int nsbi = (nrows - J0 + SUPER - 1) / SUPER;
  const int nsbj = (J1 - J0 + SUPER - 1) / SUPER;
  for (int sbi = 0; sbi < nsbi; ++sbi)
    for (int sbj = 0; sbj <= sbi && sbj < nsbj; ++sbj)
      for (int tj = 0; tj < SUPER; ++tj) {
        const int J = J0 + sbj * SUPER + tj;
        if (J >= J1) break;
        for (int ti = 0; ti < SUPER; ++ti) {
          const int I = J0 + sbi * SUPER + ti;
          if (I >= nrows) break;
          if (I >= J) out.push_back(I | (J << 16));
        }
      }
}
// border rows only: I in [nt, nrows), J in [J0, J1).  ib_lim: border tiles ib >= ib_lim are still
// exactly zero in the panel columns of this step (identity-structured right-hand sides) and are skipped.
void append_border(std::vector<int>& out, int J0, int J1, int nt, int nrows, int ib_lim) {
  for (int J = J0; J < J1; ++J)
    for (int I = nt; I < nrows && I - nt < ib_lim; ++I) out.push_back(I | (J << 16));
}

void build_plan(VcCholPlan& pl, int nt, int nbt, int nb, bool lookahead, int mode) {
  if (pl.nt == nt && pl.nbt == nbt && pl.nb == nb && pl.lookahead == lookahead && pl.mode == mode) return;
  if (nt + nbt > 0x7fff) throw VcError(VC_ERR_INVALID, "matrix too large for the packed tile index");
  std::vector<int> tiles;
  pl.calls.clear();
  const int nrows = nt + nbt;
  auto add = [&](int k0, int K, int J0, int J1) {
    VcUpdateCall c;
    c.k0 = k0; c.K = K; c.tile_off = (int64_t)tiles.size();
    if (J1 > J0) {
      if (mode == VC_SWEEP_FACTOR) append_trapezoid(tiles, J0, J1, nrows);
      else append_border(tiles, J0, J1, nt, nrows, mode == VC_SWEEP_SOLVE_IDENT ? (k0 + K) / TILE : INT_MAX);
    }
    c.ntiles = (int)((int64_t)tiles.size() - c.tile_off);
    pl.calls.push_back(c);
  };
  // must mirror the launch order of vc_chol_factor exactly
  for (int ob = 0; ob < nt; ob += nb) {
    const int oe = std::min(ob + nb, nt);
    for (int jt = ob; jt < oe; ++jt) add(jt * TILE, TILE, jt + 1, oe);
    if (oe >= nt) break;
    const int K = (oe - ob) * TILE;
    if (!lookahead) {
      add(ob * TILE, K, oe, nt);
    } else {
      const int ne = std::min(oe + nb, nt);
      add(ob * TILE, K, oe, ne);
      add(ob * TILE, K, ne, nt);
    }
  }
  if ((int64_t)tiles.size() > pl.tiles_cap) {
    if (pl.tiles_dev) cudaFree(pl.tiles_dev);
    pl.tiles_dev = nullptr;
    pl.tiles_cap = 0;
    VC_CUDA_TRY(cudaMalloc(&pl.tiles_dev, std::max<size_t>(tiles.size(), 1) * sizeof(int)));
    pl.tiles_cap = (int64_t)tiles.size();
  }
  if (!tiles.empty())
    VC_CUDA_TRY(cudaMemcpy(pl.tiles_dev, tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice));
  pl.nt = nt; pl.nbt = nbt; pl.nb = nb; pl.lookahead = lookahead; pl.mode = mode;
}

}  // namespace

int64_t vc_layout_fill(std::vector<VcColBlock>& cb, const std::vector<int>& r0, const std::vector<int>& ncols,
                       int local_rows) {
  cb.resize(r0.size());
  int64_t off = 0;
  for (size_t c = 0; c < r0.size(); ++c) {
    const int rows = std::max(local_rows - r0[c], 0);
    const int ld_tiles = std::max(rows, 1) | 1;        // odd: the column stride is never a multiple of 2 KB
    cb[c].off = off;
    cb[c].ld = ld_tiles * VC_TILE;
    cb[c].r0 = r0[c];
    off += (int64_t)cb[c].ld * ncols[c] * VC_TILE;
  }
  return std::max<int64_t>(off, 1);
}

void vc_chol_init() {
  static bool done_dev[64] = {false};
  int dev = 0;
  VC_CUDA_TRY(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return;
  VC_CUDA_TRY(cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_syrk_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_syrk_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_potrf128, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_potrf128b<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRFB_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_potrf128b<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRFB_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_potrf128c<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRFC_SMEM));
  VC_CUDA_TRY(cudaFuncSetAttribute(k_potrf128c<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRFC_SMEM));
  done = true;
}

// VC_POTRF_B=1: the previous blocked diagonal-tile kernel instead of k_potrf128c (A/B runs)
static bool potrf_use_b() {
  static const bool b = getenv("VC_POTRF_B") != nullptr;
  return b;
}

// inverse of every diagonal tile of a factor that is already resident (no factorisation)
void vc_chol_invert_diag_tiles(const VcMatrix& M, VcCholWork& w, int64_t* launches) {
  vc_launch_reset_info(w.info, w.st_main, launches);
  for (int jt = 0; jt < M.nt; ++jt) {
    int64_t ld;
    double* diag = vc_host_tile_ptr(M, jt, jt, &ld);
    const VcPotrfBatch one{1, 0, 0, 0};
    if (potrf_use_b()) k_potrf128b<false><<<1, POTRF_THREADS, POTRFB_SMEM, w.st_main>>>(diag, ld, jt, w.linv, w.logdet_part, w.info, one);
    else k_potrf128c<false><<<1, POTRF_THREADS, POTRFC_SMEM, w.st_main>>>(diag, ld, jt, w.linv, w.logdet_part, w.info, one);
    ++*launches;
  }
}

void vc_chol_free_graphs(VcCholWork& w);
void vc_chol_free_plans(VcCholWork& w) {
  vc_chol_free_graphs(w);
  if (w.plan_factor.tiles_dev) cudaFree(w.plan_factor.tiles_dev);
  if (w.plan_solve.tiles_dev) cudaFree(w.plan_solve.tiles_dev);
  w.plan_factor = VcCholPlan();
  w.plan_solve = VcCholPlan();
}

// ---------------------------------------------------------------------------------------------
// Timeline tracing (development aid): VC_TRACE=<file> brackets every factorisation kernel with CUDA
// events on its own stream and appends "kernel stream arg0 arg1 start_us end_us" lines per
// factorisation.  Off by default; the events cost ~1-2 us per launch when on.
// ---------------------------------------------------------------------------------------------
namespace {
struct TraceRec { const char* name; uintptr_t stream; int a, b; cudaEvent_t e0, e1; };
struct Tracer {
  bool on = false;
  std::string path;
  std::vector<TraceRec> recs;
  Tracer() {
    const char* p = getenv("VC_TRACE");
    if (p && *p) { on = true; path = p; }
  }
};
Tracer& tracer() { static Tracer t; return t; }
struct TraceScope {
  cudaStream_t st; bool on;
  TraceScope(const char* name, cudaStream_t s, int a, int b) : st(s), on(tracer().on) {
    if (!on) return;
    TraceRec r{name, (uintptr_t)s, a, b, nullptr, nullptr};
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, s);
    tracer().recs.push_back(r);
  }
  ~TraceScope() { if (on) cudaEventRecord(tracer().recs.back().e1, st); }
};
void trace_dump(int nt, int nb) {
  Tracer& t = tracer();
  if (!t.on || t.recs.empty()) return;
  cudaDeviceSynchronize();
  FILE* f = fopen(t.path.c_str(), "a");
  if (f) fprintf(f, "# factorisation nt=%d nb=%d launches=%zu\n", nt, nb, t.recs.size());
  for (auto& r : t.recs) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, t.recs[0].e0, r.e0);
    cudaEventElapsedTime(&b, t.recs[0].e0, r.e1);
    if (f) fprintf(f, "%s %llx %d %d %.2f %.2f\n", r.name, (unsigned long long)r.stream, r.a, r.b, a * 1e3, b * 1e3);
  }
  for (auto& r : t.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  if (f) fclose(f);
  t.recs.clear();
}
}  // namespace

void vc_launch_potrf(double* diag, int64_t lda, int jt, double* linv_all, double* logdet_part, int* info,
                     bool legacy, cudaStream_t st, int64_t* launches, VcPotrfBatch bat) {
  TraceScope ts("potrf", st, jt, bat.n);
  const int nb = std::max(bat.n, 1);
  if (legacy) k_potrf128<<<nb, POTRF_THREADS, POTRF_SMEM, st>>>(diag, lda, jt, linv_all, logdet_part, info, bat);
  else if (potrf_use_b()) k_potrf128b<true><<<nb, POTRF_THREADS, POTRFB_SMEM, st>>>(diag, lda, jt, linv_all, logdet_part, info, bat);
  else k_potrf128c<true><<<nb, POTRF_THREADS, POTRFC_SMEM, st>>>(diag, lda, jt, linv_all, logdet_part, info, bat);
  ++*launches;
}
void vc_launch_trsm(const VcGemmArgs& g, int grid, cudaStream_t st, int64_t* launches) {
  if (grid <= 0) return;
  TraceScope ts("trsm", st, g.jt, grid);
  k_trsm_panel<<<dim3(grid, std::max(g.nbatch, 1)), GEMM_THREADS, GEMM_SMEM, st>>>(g);
  ++*launches;
}
void vc_launch_syrk(const VcGemmArgs& g, int num_sms, bool legacy, cudaStream_t st, int64_t* launches) {
  if (g.ntiles <= 0) return;
  const int nbatch = std::max(g.nbatch, 1);
  if (legacy) {
    TraceScope ts(g.K == TILE ? "syrk128" : "syrk", st, g.ntiles, g.ntiles);
    k_syrk_tiles<<<dim3(g.ntiles, nbatch), GEMM_THREADS, GEMM_SMEM, st>>>(g);
    ++*launches;
    return;
  }
  const int total = g.ntiles * nbatch;       // linear index over (slot, tile)
  const int grid = std::min(total, std::max(num_sms, 1));
  // A persistent grid walks ntiles / grid tiles per CTA; a remainder of a few tiles costs a whole extra
  // tile time on an almost empty machine.  Run a small remainder as 32-row strips instead (a quarter of
  // the time, same arithmetic).  VC_NO_STRIPS=1 disables.
  static const bool no_strips = getenv("VC_NO_STRIPS") != nullptr;
  const int rem = total % grid;
  const int tail = (!no_strips && total > grid && rem > 0 && 100 * rem <= 45 * grid) ? rem : 0;
  if (tail == 0) {
    TraceScope ts(g.K == TILE ? "syrk128" : "syrk", st, total, grid);
    k_syrk_ws<<<grid, WS_THREADS, WS_SMEM, st>>>(g);
    ++*launches;
    return;
  }
  // the persistent kernel takes the linear indices [0, total - tail), the strips the last `tail` ones
  VcGemmArgs m = g;
  if (nbatch == 1) {
    m.ntiles = total - tail;
  } else {
    // with several slots the cut does not fall on a slot boundary: k_syrk_ws stops at `limit`
    m.lin_stop = total - tail;
  }
  {
    TraceScope ts(g.K == TILE ? "syrk128" : "syrk", st, total - tail, grid);
    k_syrk_ws<<<grid, WS_THREADS, WS_SMEM, st>>>(m);
    ++*launches;
  }
  VcGemmArgs t = g;
  t.lin_first = total - tail;                 // first linear index handled by the strips
  TraceScope ts("utail", st, tail, 4 * tail);
  k_strip<<<4 * tail, GEMM_THREADS, GEMM_SMEM, st>>>(t);
  ++*launches;
}
void vc_launch_reset_info(int* info, cudaStream_t st, int64_t* launches, int nbatch) {
  k_reset_info<<<std::max(nbatch, 1), 1, 0, st>>>(info);
  ++*launches;
}
void vc_launch_pack_tiles(const VcGemmArgs& g, const int* tiles, int ntiles, int lcol0, int K, double* dst,
                          int tile0, cudaStream_t st, int64_t* launches, const VcPanelMap* map) {
  if (ntiles <= 0) return;
  VcPanelMap m{};
  if (map) m = *map;
  k_pack_tiles<<<ntiles, 256, 0, st>>>(g, tiles, lcol0, K, dst, tile0, m);
  ++*launches;
}
void vc_append_trapezoid(std::vector<int>& out, int J0, int J1, int nrows) { append_trapezoid(out, J0, J1, nrows); }

static GemmArgs base_args(const VcMatrix& M) {
  GemmArgs g{};
  g.a = M.a; g.cb = M.cb; g.blk = M.blk; g.pr = 1; g.pc = 1; g.b = M.b; g.ldb = M.ldb; g.nt = M.nt;
  g.nbatch = std::max(M.nbatch, 1); g.a_es = M.a_es; g.b_es = M.b_es;
  g.linv_es = (int64_t)M.nt * TILE * TILE;
  return g;
}

static void launch_update(const VcMatrix& M, const VcCholWork& w, const VcCholPlan& pl, size_t call,
                          cudaStream_t st, int64_t* launches, int reserve_sms = 0, int first = 0, int count = -1) {
  const VcUpdateCall& c = pl.calls[call];
  if (count < 0) count = c.ntiles - first;
  if (count <= 0) return;
  GemmArgs g = base_args(M);
  g.k0 = c.k0; g.K = c.K; g.tiles = pl.tiles_dev + c.tile_off + first; g.ntiles = count;
  vc_launch_syrk(g, w.num_sms - reserve_sms, w.legacy_syrk, st, launches);
}

// A persistent update kernel keeps every SM it starts on until it is done, so the lookahead panel
// (other stream) only overlaps if some SMs are left free.  Pick the number R of SMs to leave to the
// panel so that max(panel time on R SMs, trailing update on num_sms - R SMs) is smallest.
int vc_pick_reserved_sms(int num_sms, int rows_t /*row tiles below the panel*/, int nb,
                         int rest_tiles /*per slot*/, int* tiles_beside_panel, int nbatch) {
  *tiles_beside_panel = rest_tiles;
  if (rest_tiles <= 0) return 0;
  const double t_potrf = 45.0, t_tile128 = 20.0, t_tile_nb = 17.0 * nb;   // microseconds
  const double panel_serial = nb * (t_potrf + 2 * t_tile128);
  const double panel_par = (nb + nb * (nb - 1) / 2.0) * rows_t * t_tile128 * nbatch;
  rest_tiles *= nbatch;
  int best = 1;
  double best_t = 1e300, best_tp = 0.0;
  // every slot's diagonal tile is factored by its own CTA: the potrf chain needs nbatch SMs at once
  for (int r = std::min(std::max(nbatch, 1), num_sms / 2); r <= num_sms / 2; ++r) {
    const double tp = panel_serial + panel_par / r;
    const double tu = (double)((rest_tiles + (num_sms - r) - 1) / (num_sms - r)) * t_tile_nb;
    const double t = std::max(tp, tu);
    if (t < best_t) { best_t = t; best = r; best_tp = tp; }
  }
  // Only the part of the update that overlaps the panel needs to leave SMs free: run twice the
  // panel's estimated duration on num_sms - R CTAs, the remainder on the full machine afterwards.
  const double waves = 2.0 * best_tp / t_tile_nb;
  const double tiles = waves * (num_sms - best);
  if (nbatch == 1 && tiles < 0.7 * rest_tiles) {
    int nA = (int)(tiles / (num_sms - best) + 1.0) * (num_sms - best);
    *tiles_beside_panel = std::min(nA, rest_tiles);
  }
  return best;
}
static void chol_factor_enqueue(const VcMatrix& M, VcCholWork& w, int64_t* launches, int mode) {
  const int nt = M.nt, nbt = M.nbt, nb = w.nb_tiles;
  // a K-panel of the trailing update must lie inside ONE storage block column (uniform leading dimension)
  if (nb != M.blk) throw VcError(VC_ERR_INVALID, "outer Cholesky block differs from the storage block of the matrix");
  const bool solve = (mode != VC_SWEEP_FACTOR);
  const bool la = w.lookahead && !solve;
  VcCholPlan& pl = solve ? w.plan_solve : w.plan_factor;
  const bool ident = (mode == VC_SWEEP_SOLVE_IDENT);
  build_plan(pl, nt, nbt, nb, la, mode);
  cudaStream_t sm = w.st_main, sp = la ? w.st_panel : w.st_main;
  const int nbatch = std::max(M.nbatch, 1);
  if (!solve) vc_launch_reset_info(w.info, sm, launches, nbatch);
  if (la) {
    VC_CUDA_TRY(cudaEventRecord(w.ev_update, sm));
    VC_CUDA_TRY(cudaStreamWaitEvent(sp, w.ev_update, 0));
  }
  // plan call index of the first inner update of every outer block (mirrors build_plan)
  std::vector<size_t> first_call;
  {
    size_t c = 0;
    for (int ob = 0; ob < nt; ob += nb) {
      const int oe = std::min(ob + nb, nt);
      first_call.push_back(c);
      c += (size_t)(oe - ob);
      if (oe < nt) c += la ? 2 : 1;
    }
  }
  auto panel = [&](int ob) {   // potrf / TRSM / inner update of every 128-wide sub-panel of block ob
    const int oe = std::min(ob + nb, nt);
    size_t call = first_call[ob / nb];
    for (int jt = ob; jt < oe; ++jt) {
      if (!solve) {
        int64_t ld;
        double* diag = vc_host_tile_ptr(M, jt, jt, &ld);
        vc_launch_potrf(diag, ld, jt, w.linv, w.logdet_part, w.info, w.legacy_potrf, sp, launches,
                        VcPotrfBatch{nbatch, M.a_es, (int64_t)nt * TILE * TILE, nt});
      }
      GemmArgs g = base_args(M);
      g.linv = w.linv + (int64_t)jt * TILE * TILE;
      g.k0 = jt * TILE; g.K = TILE; g.jt = jt;
      g.n_main = solve ? 0 : nt - (jt + 1);
      // identity-structured borders: tile ib is still zero in columns < 128 ib
      vc_launch_trsm(g, g.n_main + (ident ? std::min(nbt, jt + 1) : nbt), sp, launches);
      launch_update(M, w, pl, call++, sp, launches);
    }
  };
  if (!la) {
    for (int ob = 0; ob < nt; ob += nb) {
      const int oe = std::min(ob + nb, nt);
      panel(ob);
      if (oe < nt) launch_update(M, w, pl, first_call[ob / nb] + (size_t)(oe - ob), sm, launches);
    }
    trace_dump(nt, nb);
    return;
  }
  // Lookahead, software-pipelined over the outer blocks:
  //   sm:  col(s) | rest A(s) on num_sms - R CTAs ............ | rest B(s) on every SM | col(s+1) ...
  //   sp:          panel(s+1): potrf / TRSM / inner updates ---^ (rest B and col(s+1) wait for it)
  // The next panel only touches columns [oe, ne), which the rest update neither reads nor writes.
  panel(0);
  VC_CUDA_TRY(cudaEventRecord(w.ev_panel, sp));
  for (int ob = 0; ob < nt; ob += nb) {
    const int oe = std::min(ob + nb, nt);
    if (oe >= nt) break;
    const size_t c_col = first_call[ob / nb] + (size_t)(oe - ob), c_rest = c_col + 1;
    VC_CUDA_TRY(cudaStreamWaitEvent(sm, w.ev_panel, 0));
    launch_update(M, w, pl, c_col, sm, launches);
    VC_CUDA_TRY(cudaEventRecord(w.ev_col, sm));
    VC_CUDA_TRY(cudaStreamWaitEvent(sp, w.ev_col, 0));
    const int rest_tiles = pl.calls[c_rest].ntiles;
    int nA = rest_tiles;
    int reserve = w.legacy_syrk ? 0 : vc_pick_reserved_sms(w.num_sms, nt + nbt - oe, nb, rest_tiles, &nA, nbatch);
    {  // tuning overrides: VC_RESERVE_SMS = fixed R, VC_NO_SPLIT = run the whole rest update on num_sms - R
      static const int fixed = getenv("VC_RESERVE_SMS") ? atoi(getenv("VC_RESERVE_SMS")) : -1;
      static const bool nosplit = getenv("VC_NO_SPLIT") != nullptr;
      if (fixed >= 0 && !w.legacy_syrk) reserve = fixed;
      if (nosplit) nA = rest_tiles;
    }
    launch_update(M, w, pl, c_rest, sm, launches, reserve, 0, nA);
    panel(oe);
    VC_CUDA_TRY(cudaEventRecord(w.ev_panel, sp));
    if (nA < rest_tiles) {
      VC_CUDA_TRY(cudaStreamWaitEvent(sm, w.ev_panel, 0));
      launch_update(M, w, pl, c_rest, sm, launches, 0, nA, rest_tiles - nA);
    }
  }
  VC_CUDA_TRY(cudaStreamWaitEvent(sm, w.ev_panel, 0));
  trace_dump(nt, nb);
}

// CUDA graph of one factorisation (small problems only).  At n <= 8 192 a factorisation is a chain of 40-600
// short launches on two streams; replaying the captured graph takes the per-launch CPU cost and most of the
// inter-kernel gaps out of it.  Everything in the sequence is fixed for a given (matrix buffers, slot count):
// kernel arguments are addresses and tile lists, the parameters of the evaluations live in device memory (the
// build is outside the graph).  VC_GRAPH=0 disables; tracing disables.
constexpr int GRAPH_NT_MAX = 64;
void vc_chol_free_graphs(VcCholWork& w) {
  for (auto& g : w.graphs) if (g.exec) cudaGraphExecDestroy((cudaGraphExec_t)g.exec);
  w.graphs.clear();
}
void vc_chol_factor(const VcMatrix& M, VcCholWork& w, int64_t* launches, int mode) {
  static const bool graphs_on = !(getenv("VC_GRAPH") && atoi(getenv("VC_GRAPH")) == 0);
  const bool eligible = graphs_on && mode == VC_SWEEP_FACTOR && w.lookahead && !w.legacy_syrk && !tracer().on &&
                        M.nt <= GRAPH_NT_MAX;
  if (!eligible) {
    chol_factor_enqueue(M, w, launches, mode);
    return;
  }
  const int nbatch = std::max(M.nbatch, 1);
  for (auto& g : w.graphs)
    if (g.a == M.a && g.b == M.b && g.nbatch == nbatch && g.nt == M.nt && g.nbt == M.nbt) {
      VC_CUDA_TRY(cudaGraphLaunch((cudaGraphExec_t)g.exec, w.st_main));
      *launches += g.launches;
      return;
    }
  // the tile lists are uploaded with a blocking copy: that must happen before the capture starts
  build_plan(w.plan_factor, M.nt, M.nbt, w.nb_tiles, true, mode);
  const int64_t l0 = *launches;
  cudaGraph_t graph = nullptr;
  VC_CUDA_TRY(cudaStreamBeginCapture(w.st_main, cudaStreamCaptureModeThreadLocal));
  try {
    chol_factor_enqueue(M, w, launches, mode);
  } catch (...) {
    cudaStreamEndCapture(w.st_main, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  VC_CUDA_TRY(cudaStreamEndCapture(w.st_main, &graph));
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {                       // no graph to be had: run the sequence directly
    (void)cudaGetLastError();
    *launches = l0;
    chol_factor_enqueue(M, w, launches, mode);
    return;
  }
  if (w.graphs.size() >= 8) {                    // a handful of slot counts occur in practice; drop the oldest
    cudaGraphExecDestroy((cudaGraphExec_t)w.graphs.front().exec);
    w.graphs.erase(w.graphs.begin());
  }
  w.graphs.push_back(VcGraphEntry{M.a, M.b, nbatch, M.nt, M.nbt, (void*)exec, *launches - l0});
  VC_CUDA_TRY(cudaGraphLaunch(exec, w.st_main));
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
