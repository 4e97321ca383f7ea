// vc_common.cuh -- shared definitions of the sm_100a likelihood library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/varycoef_b200.h"

#define VC_TILE 128            // tile edge of every kernel (rows/cols of Sigma are padded to it)
#define VC_MAXQ 32             // SVCs per model supported by the by-value kernel parameter block
#define VC_MAXD 8              // coordinate columns supported
#define VC_BORDER_ROWS 128     // one extra row tile below Sigma holds [X y]^T (p + 1 <= 128)

// covariance parameters of one evaluation, passed to kernels by value (no H2D copy on the path)
struct VcTheta {
  int q;
  int cov_id;
  int taper_id;
  int dist_method;
  double mink_p;
  double inv_delta;            // 1 / taper range
  double tau2;                 // nugget variance
  double inv_range[VC_MAXQ];   // 1 / rho_k
  double sill[VC_MAXQ];        // sigma_k^2
};

// layout of the per-evaluation result record in device memory (doubles)
#define VC_RES_N2LL 0
#define VC_RES_LOGDET 1
#define VC_RES_QUAD 2
#define VC_RES_INFO 3
#define VC_RES_MU 4            // p doubles follow
#define VC_RES_STRIDE (4 + 128)

#define VC_CUDA_TRY(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      char _b[512];                                                                    \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
               __FILE__, __LINE__);                                                    \
      throw VcError(_e == cudaErrorMemoryAllocation ? VC_ERR_OOM : VC_ERR_CUDA, _b);   \
    }                                                                                  \
  } while (0)

struct VcError {
  int code;
  std::string msg;
  VcError(int c, const std::string& m) : code(c), msg(m) {}
};

// ---- storage of the lower triangle ----------------------------------------------------------
// Only the lower trapezoid of every BLOCK COLUMN is stored (blk tile columns wide; blk = the outer
// Cholesky block = the distribution block of the sharded path): block column c holds the row tiles
// [r0, rows) of its columns, column-major with its own leading dimension (an odd number of 128-double
// tiles, so that no column stride is a multiple of 2 KB).  Tile (lI, lJ) in LOCAL tile indices:
//   a + cb[lJ / blk].off + (lI - r0) * 128 + ((lJ % blk) * 128 + col) * ld
// n = 50 000: 10.3 GB instead of 20.1 GB for the full square; n = 150 000: 91 GB -- fits one B200.
struct VcColBlock {
  int64_t off;           // doubles from the start of the allocation to (row tile r0, first column)
  int32_t ld;            // leading dimension in doubles
  int32_t r0;            // first LOCAL row tile stored in this block column
};
#if defined(__CUDACC__)
__device__ __forceinline__ double* vc_tile_ptr(double* a, const VcColBlock* __restrict__ cb, int blk, int lI, int lJ,
                                               int64_t& ld) {
  const int c = lJ / blk;
  const longlong2 e0 = __ldg(reinterpret_cast<const longlong2*>(cb + c));   // one 16-byte load: off | (ld, r0)
  const int eld = (int)(e0.y & 0xffffffffll), er0 = (int)(e0.y >> 32);
  ld = eld;
  return a + e0.x + (int64_t)(lI - er0) * VC_TILE + (int64_t)(lJ - c * blk) * VC_TILE * (int64_t)eld;
}
// element (i, j), i >= j, of a NON-distributed matrix
__device__ __forceinline__ double* vc_elem_ptr(double* a, const VcColBlock* __restrict__ cb, int blk, int64_t i, int64_t j) {
  int64_t ld;
  double* t = vc_tile_ptr(a, cb, blk, (int)(i >> 7), (int)(j >> 7), ld);
  return t + (i & (VC_TILE - 1)) + (j & (VC_TILE - 1)) * ld;
}
#endif
// fills `cb` for block columns whose first stored local row tile is r0[c]; local_rows = row tiles held.
// Returns the number of doubles to allocate.  ncols[c] = tile columns of block column c (<= blk).
int64_t vc_layout_fill(std::vector<VcColBlock>& cb, const std::vector<int>& r0, const std::vector<int>& ncols,
                       int local_rows);

// ---- launchers implemented in the .cu files ------------------------------------------------
struct VcMatrix {        // padded Sigma / factor in HBM (lower block-column trapezoids), plus its border rows
  double* a;
  const VcColBlock* cb;       // device table, one entry per local block column
  const VcColBlock* cb_host;  // the same table on the host (owned by the handle)
  int blk;               // tile columns per block column
  int64_t a_doubles;     // doubles of ONE evaluation slot behind `a`
  int nbatch;            // evaluation slots swept by one launch (batched small-n path); slot e lives at
  int64_t a_es, b_es;    //   a + e * a_es, b + e * b_es
  int64_t n;             // real order
  int64_t n_pad;         // multiple of VC_TILE
  int nt;                // n_pad / VC_TILE  (column tiles)
  double* b;             // border: (128 * nbt) x n_pad column-major; rows = right-hand sides B^T,
  int64_t ldb;           //         turned into (L^-1 B)^T by the factorisation / solve sweep
  int nbt;               // border row tiles in use
};

// vc_build.cu
// th_dev != nullptr: batched build, slot e (grid.y) uses th_dev[e]; th then only supplies the model (cov, taper, q)
void vc_launch_build(const VcMatrix& M, const VcTheta& th, int d, const double* locs_pad,
                     const double* w_pad, cudaStream_t st, int64_t* launches, const VcTheta* th_dev = nullptr);
void vc_launch_border(const VcMatrix& M, int p, const double* x_pad, const double* y_pad,
                      cudaStream_t st, int64_t* launches);
void vc_launch_sigma_full(double* out, int64_t n, int64_t n_pad, const VcTheta& th, int d,
                          const double* locs_pad, const double* w_pad, cudaStream_t st,
                          int64_t* launches);
void vc_build_init();
void vc_launch_build_tiles(const VcMatrix& M, const VcTheta& th, int d,
                           const double* locs_pad, const double* w_pad, const int* tiles, int ntiles,
                           int pr, int pc, cudaStream_t st, int64_t* launches);
void vc_launch_border_dist(double* b, int64_t ldb, int64_t n_pad, int p, const double* x_pad,
                           const double* y_pad, const int* col_tiles_dev, int nloc, cudaStream_t st,
                           int64_t* launches);

// vc_chol.cu
// Order of the row tiles inside a gathered panel buffer of the sharded path: OWNER-major -- the tiles of
// process row 0 first (then its border tiles), then process row 1, ... -- so that every owner's part is one
// contiguous range and travels in ONE broadcast (round 1: tile order, one broadcast per distribution block,
// 147 latency-bound NCCL operations per outer block at n = 150 000).
struct VcPanelMap {
  int pr;                 // process rows; 0 = plain tile order (position = I - tile0)
  int nb, nt;             // tiles per distribution block, main row tiles
  int main0;              // main tiles of process row 0 in the buffer (its border tiles follow them)
  int chunk_off[8];       // first position of every process row's part
  int bfirst[8];          // first distribution block of every process row that is in the buffer
};
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
int vc_panel_pos(const VcPanelMap& m, int I, int tile0) {
  if (m.pr == 0) return I - tile0;
  if (I >= m.nt) return m.chunk_off[0] + m.main0 + (I - m.nt);
  const int B = I / m.nb, r = B % m.pr;
  return m.chunk_off[r] + ((B - m.bfirst[r]) / m.pr) * m.nb + (I - B * m.nb);
}
struct VcOpSrc {          // where a GEMM operand row tile comes from
  const double* pnl;      // gathered panel buffer, or nullptr = the matrix itself at column k0
  int64_t ld;             // leading dimension inside a tile of the buffer
  int64_t tile_stride;    // doubles between consecutive row tiles of the buffer
  int tile0;              // global row tile stored at the start of the buffer (plain order)
  int col0;               // first panel column inside the buffer
  VcPanelMap map;         // position of a row tile in the buffer
};
struct VcGemmArgs {       // by-value argument block of k_trsm_panel / k_syrk_ws / k_syrk_tiles
  double* a;              // (local) matrix, row tiles 0 .. nt-1 (global numbering)
  const VcColBlock* cb;   // its block-column table
  int nbatch;             // evaluation slots handled by this launch (grid.y, or the persistent tile index / ntiles)
  int64_t a_es, b_es, linv_es;   // doubles between consecutive slots of a / b / linv
  double* b;              // border rows, row tiles nt .. nt+nbt-1
  int64_t ldb;
  int nt;
  int blk, pr, pc;        // tiles per (storage = distribution) block; process grid (1 x 1 = not distributed)
  const double* linv;     // TRSM: 128x128 inverse factor (ld 128)
  int k0;                 // first panel column in LOCAL elements (operands taken from the matrix)
  int K;                  // panel depth (multiple of 16)
  int jt;                 // TRSM: panel tile column (global)
  int n_main;             // TRSM without tile list: main row tiles handled
  const int* tiles;       // packed (I | J << 16), global tile indices
  int ntiles;
  VcOpSrc src_a, src_b;   // operand sources of the update kernels
  int i_skip;             // TRSM without tile list: main row tiles start at jt + 1 + i_skip
  int lin_first, lin_stop;  // update kernels: first linear (slot, tile) index of k_strip / stop index of k_syrk_ws (0 = all)
};
struct VcUpdateCall {     // one trailing / inner update of the factorisation plan
  int k0, K;             // panel columns [k0, k0 + K)
  int64_t tile_off;      // into the device tile list
  int ntiles;
};
struct VcCholPlan {       // tile lists of every update launch, precomputed per (nt, nbt, nb, mode)
  int nt = -1, nbt = -1, nb = -1, mode = -1;
  bool lookahead = false;
  std::vector<VcUpdateCall> calls;
  int* tiles_dev = nullptr;          // packed (I | J << 16), L2-friendly super-block order
  int64_t tiles_cap = 0;
};
struct VcGraphEntry {     // an instantiated CUDA graph of one factorisation (vc_chol_factor)
  const double* a; const double* b; int nbatch, nt, nbt;
  void* exec;             // cudaGraphExec_t
  int64_t launches;       // kernels it replays
};
struct VcCholWork {
  double* linv;          // per slot: nt x (128 x 128), the inverse of every diagonal factor tile
  double* logdet_part;   // per slot: nt partial sums of log L_jj
  int* info;             // per slot: first failing leading minor (INT_MAX if none)
  int nb_tiles;          // outer block in tiles (nb_outer / 128)
  bool lookahead;
  bool legacy_syrk;      // use the non-persistent cp.async SYRK kernel (A/B testing)
  bool legacy_potrf;     // use the unblocked diagonal-tile kernel (A/B testing)
  int num_sms;
  cudaStream_t st_main, st_panel;
  cudaEvent_t ev_panel, ev_update, ev_col;
  VcCholPlan plan_factor, plan_solve;
  std::vector<VcGraphEntry> graphs;
};
enum { VC_SWEEP_FACTOR = 0,   // potrf + TRSM + updates on every row tile (main + border)
       VC_SWEEP_SOLVE = 1,    // factor already resident: only the border rows are swept
       VC_SWEEP_SOLVE_IDENT = 2 };  // same, border = identity rows: tiles known to be zero are skipped
void vc_chol_init();
void vc_chol_factor(const VcMatrix& M, VcCholWork& w, int64_t* launches, int mode = VC_SWEEP_FACTOR);
void vc_chol_free_plans(VcCholWork& w);
void vc_chol_free_graphs(VcCholWork& w);   // must be called when the matrix buffers are reallocated
void vc_chol_invert_diag_tiles(const VcMatrix& M, VcCholWork& w, int64_t* launches);
// raw launchers (used by the distributed driver, vc_dist.cu)
// host-side address of local tile (lI, lJ)
inline double* vc_host_tile_ptr(const VcMatrix& M, int lI, int lJ, int64_t* ld) {
  const VcColBlock& e = M.cb_host[lJ / M.blk];
  *ld = e.ld;
  return M.a + e.off + (int64_t)(lI - e.r0) * VC_TILE + (int64_t)(lJ % M.blk) * VC_TILE * (int64_t)e.ld;
}
struct VcPotrfBatch { int n; int64_t a_es, linv_es; int nt; };   // slot e: diag + e a_es, linv + e linv_es, logdet + e nt, info + e
void vc_launch_potrf(double* diag, int64_t lda, int jt, double* linv_all, double* logdet_part, int* info,
                     bool legacy, cudaStream_t st, int64_t* launches, VcPotrfBatch bat = VcPotrfBatch{1, 0, 0, 0});
void vc_launch_trsm(const VcGemmArgs& g, int grid, cudaStream_t st, int64_t* launches);
void vc_launch_syrk(const VcGemmArgs& g, int num_sms, bool legacy, cudaStream_t st, int64_t* launches);
void vc_launch_reset_info(int* info, cudaStream_t st, int64_t* launches, int nbatch = 1);
void vc_launch_pack_tiles(const VcGemmArgs& g, const int* tiles, int ntiles, int lcol0, int K, double* dst,
                          int tile0, cudaStream_t st, int64_t* launches, const VcPanelMap* map = nullptr);
void vc_append_trapezoid(std::vector<int>& out, int J0, int J1, int nrows);
int vc_pick_reserved_sms(int num_sms, int rows_t, int nb, int rest_tiles, int* tiles_beside_panel, int nbatch = 1);

// vc_finalize.cu
struct VcFinalWork {
  double* gram_part;     // blocks x npairs x 2 (hi, lo)
  double* quad_part;     // blocks x 2
  int blocks;
};
// njobs finalize jobs in one launch: job j reads the whitened border / log-det / info of slot job_slot_dev[j]
// (nullptr: slot 0) and mu_in_dev + 128 j, writes result_dev + VC_RES_STRIDE j; fw holds njobs partial buffers;
// gram_out_dev receives the Gram matrix of slot 0
void vc_launch_finalize(const VcMatrix& M, int p, int mu_mode, const double* mu_in_dev,
                        const double* logdet_part, const int* info, VcFinalWork& fw,
                        double* result_dev, cudaStream_t st, int64_t* launches,
                        double* gram_out_dev = nullptr, int njobs = 1, const int* job_slot_dev = nullptr);
void vc_launch_gram_local(const double* b, int64_t ldb, int64_t ncols, int p, VcFinalWork& fw, double* gram_dd,
                          cudaStream_t st, int64_t* launches);
void vc_launch_sum(const double* v, int n, double* out, cudaStream_t st, int64_t* launches);
void vc_launch_solve_reduced(const double* gram_dd, int p, int mu_mode, const double* mu_in_dev,
                             const double* logdet_sum, const int* info, double* result_dev, double* gram_out,
                             cudaStream_t st, int64_t* launches);
void vc_launch_quad_local(const double* b, int64_t ldb, int64_t ncols, int p, const double* result_dev,
                          VcFinalWork& fw, double* quad_dd, cudaStream_t st, int64_t* launches);
void vc_launch_assemble_reduced(const double* quad_dd, int64_t n, double* result_dev, cudaStream_t st,
                                int64_t* launches);
void vc_launch_gather_whitened(const VcMatrix& M, int p, double* z_out_dev, cudaStream_t st,
                               int64_t* launches);
void vc_launch_extract_chol(const VcMatrix& M, double* r_out_dev, cudaStream_t st,
                            int64_t* launches);
