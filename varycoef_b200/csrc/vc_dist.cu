// vc_dist.cu -- K4: one -2logL evaluation sharded over a pr x pc grid of GPUs (one process per GPU).
//
// Sigma_y is built and factored 2D block-cyclic (distribution block = the outer Cholesky block,
// nb tiles of 128): tile (I, J) lives on process ((I/nb) % pr, (J/nb) % pc) at local tile
// ((I/nb)/pr * nb + I%nb, (J/nb)/pc * nb + J%nb).  The reference has nothing like it (single R
// process; optimParallel PSOCK workers, R-pkg/R/SVC_mle.R:164-200): this exists because the fp64
// Sigma_y of n = 150 000 (180 GB) does not fit one GPU.
//
// Per outer block [ob, oe):
//   1. the owner of the diagonal block factors it locally (k_potrf128b / k_trsm_panel / k_syrk_ws
//      on its own tiles) and broadcasts L_DD and the nb diagonal-tile inverses (NCCL, 2.5 MB);
//   2. the ranks of the panel's process column solve their rows against L_DD (TRSM + inner update
//      per 128-wide sub-panel, operands from the broadcast buffer);
//   3. they pack their part of the panel tile-major and every distribution block of it is
//      broadcast by its owner to all ranks (grouped ncclBroadcast over NVLink/NVSwitch);
//   4. every rank applies the trailing update to the tiles it owns, reading both operands from
//      the gathered panel buffer.
// The build is communication-free (each rank evaluates exactly its tiles from replicated O(n)
// inputs); the finalize needs three tiny all-reduces (Gram, log-det + info, quadratic form).
#include "vc_handle.cuh"
#include <dlfcn.h>
#include <limits.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <stdlib.h>

// ---- host-only block-cyclic maps (exported; unit-tested on the CPU) --------------------------
extern "C" int vc_bc_owner(int32_t bi, int32_t bj, int32_t pr, int32_t pc) {
  if (pr < 1 || pc < 1 || bi < 0 || bj < 0) return -1;
  return (bi % pr) * pc + (bj % pc);          // row-major process grid
}
extern "C" int vc_bc_local(int32_t b, int32_t nprocs_dim) {
  if (nprocs_dim < 1 || b < 0) return -1;
  return b / nprocs_dim;
}
extern "C" int vc_bc_count(int32_t nblocks, int32_t coord, int32_t nprocs_dim) {
  if (nprocs_dim < 1 || coord < 0 || coord >= nprocs_dim || nblocks < 0) return -1;
  return (nblocks - coord + nprocs_dim - 1) / nprocs_dim;
}

// trailing-update tiles (packed I | J << 16) that `rank` owns after outer block [.., oe):
// J in [oe, nt), I in [J, nt + nbt); border tiles (I >= nt) belong to process row 0.
static void trailing_tiles(int nt, int nbt, int nb, int pr, int pc, int rank, int oe, std::vector<int>& out) {
  std::vector<int> all;
  vc_append_trapezoid(all, oe, nt, nt + nbt);
  const int my_pr = rank / pc, my_pc = rank % pc;
  for (int t : all) {
    const int I = t & 0xffff, J = t >> 16;
    const int orow = (I < nt) ? (I / nb) % pr : 0;
    if (orow == my_pr && (J / nb) % pc == my_pc) out.push_back(t);
  }
}
extern "C" int vc_bc_trailing_tiles(int32_t nt, int32_t nbt, int32_t nb, int32_t pr, int32_t pc, int32_t rank,
                                    int32_t oe, int32_t* out, int32_t cap) {
  if (nt < 1 || nbt < 0 || nb < 1 || pr < 1 || pc < 1 || rank < 0 || rank >= pr * pc) return -1;
  std::vector<int> v;
  trailing_tiles(nt, nbt, nb, pr, pc, rank, oe, v);
  if (out) for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}

// Owner-major order of the row tiles >= oe in a gathered panel buffer (VcPanelMap, vc_common.cuh): the tiles of process
// row 0 (then the border tiles, which belong to it), then process row 1, ...  cnt[r] = tiles of process row r.
static void fill_panel_map(VcPanelMap& m, int cnt[8], int nt, int nbt, int nb, int pr, int oe) {
  memset(&m, 0, sizeof(m));
  m.pr = pr; m.nb = nb; m.nt = nt;
  const int nblocks = (nt + nb - 1) / nb;
  const int B0 = (oe + nb - 1) / nb;                   // first distribution block in the buffer (oe is block-aligned)
  int pos = 0;
  for (int r = 0; r < pr; ++r) {
    int bf = B0;
    while (bf % pr != r) ++bf;
    m.bfirst[r] = bf;
    m.chunk_off[r] = pos;
    int c = 0;
    for (int B = bf; B < nblocks; B += pr) c += std::min(nb, nt - B * nb);
    if (r == 0) { m.main0 = c; c += nbt; }
    cnt[r] = c;
    pos += c;
  }
}
// host-only view of that order (unit tests): per process row its first position and tile count, per row tile
// I = oe .. nt + nbt - 1 its position in the buffer.  Returns the number of tiles in the buffer.
extern "C" int vc_bc_panel_layout(int32_t nt, int32_t nbt, int32_t nb, int32_t pr, int32_t oe, int32_t* chunk_off,
                                  int32_t* chunk_cnt, int32_t* pos) {
  if (nt < 1 || nbt < 0 || nb < 1 || pr < 1 || pr > 8 || oe < 0 || oe > nt) return -1;
  VcPanelMap m;
  int cnt[8];
  fill_panel_map(m, cnt, nt, nbt, nb, pr, oe);
  int total = 0;
  for (int r = 0; r < pr; ++r) {
    if (chunk_off) chunk_off[r] = m.chunk_off[r];
    if (chunk_cnt) chunk_cnt[r] = cnt[r];
    total += cnt[r];
  }
  if (pos) for (int I = oe; I < nt + nbt; ++I) pos[I - oe] = vc_panel_pos(m, I, oe);
  return total;
}

// ---- NCCL through dlopen (the library already loaded by the host process is reused) ----------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt32_ = 2, ncclFloat64_ = 8 };
enum { ncclSum_ = 0, ncclMin_ = 3 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, void*) = nullptr;   // optional (NCCL >= 2.14)
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define VC_SYM(field, name)                                                   \
    *(void**)(&field) = dlsym(lib, name);                                     \
    if (!field) { err = std::string("libnccl lacks ") + name; return false; }
    VC_SYM(GetUniqueId, "ncclGetUniqueId")
    VC_SYM(CommInitRank, "ncclCommInitRank")
    VC_SYM(CommDestroy, "ncclCommDestroy")
    VC_SYM(Broadcast, "ncclBroadcast")
    VC_SYM(AllReduce, "ncclAllReduce")
    VC_SYM(GroupStart, "ncclGroupStart")
    VC_SYM(GroupEnd, "ncclGroupEnd")
    VC_SYM(GetErrorString, "ncclGetErrorString")
#undef VC_SYM
    *(void**)(&CommInitRankConfig) = dlsym(lib, "ncclCommInitRankConfig");
    return true;
  }
};
static NcclApi g_nccl;

#define VC_NCCL_TRY(expr)                                                                       \
  do {                                                                                          \
    int _r = (expr);                                                                            \
    if (_r != 0) {                                                                              \
      char _b[512];                                                                             \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
      throw VcError(VC_ERR_CUDA, _b);                                                           \
    }                                                                                           \
  } while (0)

extern "C" int vc_nccl_unique_id(char id_out[128]) {
  if (!id_out) return VC_ERR_INVALID;
  if (!g_nccl.load()) return vc_fail(nullptr, VC_ERR_CUDA, g_nccl.err);
  ncclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return vc_fail(nullptr, VC_ERR_CUDA, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
  memcpy(id_out, id.internal, 128);
  return VC_OK;
}

// SMs the rest update leaves free for NCCL's kernels = CTAs the communicator may use (VC_DIST_RESERVE overrides)
static int vc_dist_nccl_ctas() {
  static const int v = getenv("VC_DIST_RESERVE") ? atoi(getenv("VC_DIST_RESERVE")) : 8;
  return v;
}

// ---- the distributed state ---------------------------------------------------------------------
struct Span { int64_t off; int n; };
struct BcastOp { int64_t off; int64_t count; int root; };
struct DistStep {
  int ob, oe, diag_owner, pcp;
  // diagonal owner only
  std::vector<Span> d_trsm, d_inner;     // per sub-panel
  Span d_pack;
  // panel-column ranks
  Span rows;                             // local row tiles >= oe (+ border), packed I
  std::vector<Span> inner;               // per sub-panel: (I in rows) x (J in (jt, oe))
  std::vector<BcastOp> bcasts;           // one per process row: its part of the panel, in doubles relative to pnl
  VcPanelMap map;                        // owner-major order of the row tiles >= oe in the panel buffer
  Span trailing_cold, trailing_col, trailing_rest;   // next diagonal block, the rest of the next block's columns, the rest
};
struct VcDist {
  int rank = 0, world = 1, pr = 1, pc = 1, my_pr = 0, my_pc = 0;
  ncclComm_t comm = nullptr;
  int nt = 0, nbt = 1, nb = 4;
  int lrt = 0, lct = 0;
  std::vector<int> col_tiles;
  int* col_tiles_dev = nullptr;
  int* tiles_dev = nullptr;
  Span build;
  std::vector<DistStep> steps;
  double* dd = nullptr;       // nb tiles x (128 x nb*128) + nb x (128 x 128) inverses
  double* pnl[2] = {nullptr, nullptr};   // double-buffered gathered panel (lookahead)
  cudaEvent_t ev_dd = nullptr, ev_rows = nullptr;   // L_DD has arrived / the panel rows are packed
  double rx_bytes = 0.0;      // bytes this rank receives over NVLink per evaluation (panel + L_DD broadcasts)
  double* red = nullptr;      // [gram pairs*2 | logdet | quad(2)]
  int* info_red = nullptr;
};

static inline int loc_tile_h(int t, int nb, int P) { return ((t / nb) / P) * nb + t % nb; }

void vc_dist_destroy(vc_handle* h) {
  VcDist* D = (VcDist*)h->dist;
  if (!D) return;
  if (D->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(D->comm);
  if (D->ev_dd) cudaEventDestroy(D->ev_dd);
  if (D->ev_rows) cudaEventDestroy(D->ev_rows);
  cudaFree(D->col_tiles_dev); cudaFree(D->tiles_dev); cudaFree(D->dd); cudaFree(D->pnl[0]); cudaFree(D->pnl[1]);
  cudaFree(D->red); cudaFree(D->info_red);
  delete D;
  h->dist = nullptr;
}

static void build_dist_plan(vc_handle* h, VcDist* D) {
  const int nt = D->nt, nbt = D->nbt, nb = D->nb, pr = D->pr, pc = D->pc;
  const int nblocks = (nt + nb - 1) / nb;
  std::vector<int> tiles;
  auto span_begin = [&]() { return (int64_t)tiles.size(); };
  auto span_end = [&](int64_t off) { Span s; s.off = off; s.n = (int)((int64_t)tiles.size() - off); return s; };
  auto own_row = [&](int I) { return (I < nt ? (I / nb) % pr : 0) == D->my_pr; };
  auto own_col = [&](int J) { return (J / nb) % pc == D->my_pc; };
  // build list: local lower tiles (diagonal tiles in full)
  {
    const int64_t off = span_begin();
    for (int J = 0; J < nt; ++J)
      if (own_col(J))
        for (int I = J; I < nt; ++I)
          if (own_row(I)) tiles.push_back(I | (J << 16));
    D->build = span_end(off);
  }
  for (int b = 0; b < nblocks; ++b) {
    DistStep st;
    st.ob = b * nb;
    st.oe = std::min(st.ob + nb, nt);
    st.pcp = b % pc;
    st.diag_owner = (b % pr) * pc + st.pcp;
    const int Kb = (st.oe - st.ob) * VC_TILE;
    if (D->rank == st.diag_owner) {
      for (int jt = st.ob; jt < st.oe; ++jt) {
        int64_t off = span_begin();
        for (int I = jt + 1; I < st.oe; ++I) tiles.push_back(I | (jt << 16));
        st.d_trsm.push_back(span_end(off));
        off = span_begin();
        for (int J = jt + 1; J < st.oe; ++J)
          for (int I = J; I < st.oe; ++I) tiles.push_back(I | (J << 16));
        st.d_inner.push_back(span_end(off));
      }
      const int64_t off = span_begin();
      for (int I = st.ob; I < st.oe; ++I) tiles.push_back(I);
      st.d_pack = span_end(off);
    }
    if (D->my_pc == st.pcp) {
      int64_t off = span_begin();
      for (int I = st.oe; I < nt + nbt; ++I)
        if (own_row(I)) tiles.push_back(I | (st.ob << 16));
      st.rows = span_end(off);
      std::vector<int> rows(tiles.begin() + st.rows.off, tiles.end());
      for (int jt = st.ob; jt < st.oe; ++jt) {
        off = span_begin();
        for (int J = jt + 1; J < st.oe; ++J)
          for (int I : rows) tiles.push_back((I & 0xffff) | (J << 16));
        st.inner.push_back(span_end(off));
      }
    }
    // the gathered panel buffer holds the row tiles >= oe in OWNER-major order (VcPanelMap); every process row's
    // part is contiguous and is broadcast by its owner in the panel's process column in one operation
    {
      int cnt[8];
      fill_panel_map(st.map, cnt, nt, nbt, nb, pr, st.oe);
      for (int r = 0; r < pr; ++r)
        if (st.oe < nt && cnt[r] > 0) {
          BcastOp op;
          op.off = (int64_t)st.map.chunk_off[r] * VC_TILE * Kb;
          op.count = (int64_t)cnt[r] * VC_TILE * Kb;
          op.root = r * pc + st.pcp;
          st.bcasts.push_back(op);
        }
    }
    {
      std::vector<int> tr;
      if (st.oe < nt) trailing_tiles(nt, nbt, nb, pr, pc, D->rank, st.oe, tr);
      const int ne = std::min(st.oe + nb, nt);
      int64_t off = span_begin();
      for (int t : tr) if ((t >> 16) < ne && (t & 0xffff) < ne) tiles.push_back(t);    // the next diagonal block
      st.trailing_cold = span_end(off);
      off = span_begin();
      for (int t : tr) if ((t >> 16) < ne && (t & 0xffff) >= ne) tiles.push_back(t);
      st.trailing_col = span_end(off);
      off = span_begin();
      for (int t : tr) if ((t >> 16) >= ne) tiles.push_back(t);
      st.trailing_rest = span_end(off);
    }
    // NVLink accounting: every broadcast this rank is not the root of delivers its payload here
    if (D->world > 1) {
      if (D->rank != st.diag_owner) D->rx_bytes += 8.0 * ((double)nb * VC_TILE * nb * VC_TILE + (double)nb * VC_TILE * VC_TILE);
      for (const BcastOp& op : st.bcasts) if (op.root != D->rank) D->rx_bytes += 8.0 * (double)op.count;
    }
    D->steps.push_back(st);
  }
  VC_CUDA_TRY(cudaMalloc(&D->tiles_dev, std::max<size_t>(tiles.size(), 1) * sizeof(int)));
  if (!tiles.empty())
    VC_CUDA_TRY(cudaMemcpy(D->tiles_dev, tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice));
  (void)h;
}

extern "C" int vc_create_dist(vc_handle** out, const vc_config* cfg_in, int64_t n, int32_t d, int32_t p,
                              int32_t q, const double* locs, const double* X, const double* W,
                              const double* y, int32_t rank, int32_t world, int32_t pr, int32_t pc,
                              const char nccl_id[128]) {
  if (!out) return VC_ERR_INVALID;
  *out = nullptr;
  vc_config cfg;
  if (cfg_in) cfg = *cfg_in; else vc_config_default(&cfg);
  if (!locs || !X || !W || !y || !nccl_id) return vc_fail(nullptr, VC_ERR_INVALID, "NULL input pointer");
  if (pr < 1 || pc < 1 || pr * pc != world || rank < 0 || rank >= world)
    return vc_fail(nullptr, VC_ERR_INVALID, "process grid pr x pc must equal the world size");
  if (pr > 8) return vc_fail(nullptr, VC_ERR_INVALID, "at most 8 process rows (VcPanelMap)");
  std::string why;
  if (vc_validate_cfg(&cfg, n, d, p, q, why) != VC_OK) return vc_fail(nullptr, VC_ERR_INVALID, why);
  if (!g_nccl.load()) return vc_fail(nullptr, VC_ERR_CUDA, g_nccl.err);
  vc_handle* h = nullptr;
  VcDist* D = nullptr;
  try {
    h = vc_alloc_common(&cfg, n, d, p, q, locs, X, W, y, /*alloc_matrix=*/false);
    D = new VcDist();
    h->dist = D;
    D->rank = rank; D->world = world; D->pr = pr; D->pc = pc;
    D->my_pr = rank / pc; D->my_pc = rank % pc;
    D->nt = h->M.nt; D->nbt = 1; D->nb = h->cw.nb_tiles;
    const int nb = D->nb, nt = D->nt;
    const int nblocks = (nt + nb - 1) / nb;
    // local shapes
    int lrt = 0, lct = 0;
    for (int b = 0; b < nblocks; ++b) {
      const int tiles_in = std::min(nb, nt - b * nb);
      if (b % pr == D->my_pr) lrt += tiles_in;
      if (b % pc == D->my_pc) {
        for (int t = 0; t < tiles_in; ++t) D->col_tiles.push_back(b * nb + t);
        lct += tiles_in;
      }
    }
    D->lrt = lrt; D->lct = lct;
    // local storage: only the lower trapezoid of every local block column.  Local block column lc is global
    // block Bj = lc * pc + my_pc; the first local row block with global index >= Bj is ceil((Bj - my_pr) / pr).
    {
      std::vector<int> r0, ncols;
      for (int b = D->my_pc; b < nblocks; b += pc) {
        const int first = b <= D->my_pr ? 0 : (b - D->my_pr + pr - 1) / pr;
        r0.push_back(std::min(first * nb, lrt));
        ncols.push_back(std::min(nb, nt - b * nb));
      }
      vc_set_layout(h, r0, ncols, lrt);
    }
    h->M.ldb = VC_BORDER_ROWS;
    h->M.nbt = 1;
    VC_CUDA_TRY(cudaMalloc(&h->M.a, (size_t)h->M.a_doubles * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&h->M.b, (size_t)h->M.ldb * std::max(lct, 1) * VC_TILE * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&h->cw.logdet_part, (size_t)nt * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&D->col_tiles_dev, std::max<size_t>(D->col_tiles.size(), 1) * sizeof(int)));
    if (!D->col_tiles.empty())
      VC_CUDA_TRY(cudaMemcpy(D->col_tiles_dev, D->col_tiles.data(), D->col_tiles.size() * sizeof(int),
                             cudaMemcpyHostToDevice));
    const int64_t Kmax = (int64_t)nb * VC_TILE;
    VC_CUDA_TRY(cudaMalloc(&D->dd, ((size_t)nb * VC_TILE * Kmax + (size_t)nb * VC_TILE * VC_TILE) * sizeof(double)));
    // inverse slots: the diagonal-tile kernel writes only their lower 16x16 blocks, the rest stays zero
    VC_CUDA_TRY(cudaMemsetAsync(D->dd, 0, ((size_t)nb * VC_TILE * Kmax + (size_t)nb * VC_TILE * VC_TILE) * sizeof(double), h->st_main));
    VC_CUDA_TRY(cudaStreamSynchronize(h->st_main));
    for (int i = 0; i < 2; ++i)
      VC_CUDA_TRY(cudaMalloc(&D->pnl[i], (size_t)(nt + 1) * VC_TILE * Kmax * sizeof(double)));
    const int m = p + 1;
    VC_CUDA_TRY(cudaMalloc(&D->red, ((size_t)m * (m + 1) + 8) * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&D->info_red, sizeof(int)));
    build_dist_plan(h, D);
    ncclUniqueId id;
    memcpy(id.internal, nccl_id, 128);
    // The panel broadcasts run BESIDE the trailing update on a handful of SMs left free for them.  NCCL launches one
    // CTA per channel and a CTA that finds no free SM waits for one -- behind a persistent update kernel that means
    // until that kernel ends, which serialises broadcast and update.  So cap the communicator's CTAs at the number
    // of SMs the update leaves free (ncclConfig_t.maxCTAs, v2.17 layout of the struct; older NCCL: plain init).
    struct { size_t size; unsigned magic, version; int blocking, cgaClusterSize, minCTAs, maxCTAs; const char* netName; } ncfg =
        {0, 0xcafebeefu, 21701u, INT_MIN, INT_MIN, INT_MIN, INT_MIN, nullptr};
    ncfg.size = sizeof(ncfg);
    ncfg.maxCTAs = vc_dist_nccl_ctas();
    int nrc = -1;
    if (g_nccl.CommInitRankConfig && ncfg.maxCTAs > 0) nrc = g_nccl.CommInitRankConfig(&D->comm, world, id, rank, &ncfg);
    if (nrc != 0) VC_NCCL_TRY(g_nccl.CommInitRank(&D->comm, world, id, rank));
  } catch (const VcError& e) {
    if (h) vc_free_handle(h);
    return vc_fail(nullptr, e.code, e.msg);
  }
  *out = h;
  return VC_OK;
}

// bytes received over NVLink by this rank per evaluation, and the bound n^2 8 (1/pr + 1/pc) / 2 of a
// row / column-communicator scheme for comparison
extern "C" int vc_dist_traffic(vc_handle* h, double* rx_bytes, double* rowcol_bound_bytes) {
  if (!h || !h->dist) return VC_ERR_INVALID;
  VcDist* D = (VcDist*)h->dist;
  if (rx_bytes) *rx_bytes = D->rx_bytes;
  if (rowcol_bound_bytes)
    *rowcol_bound_bytes = D->world > 1 ? 8.0 * (double)h->n_pad * (double)h->n_pad * (1.0 / D->pr + 1.0 / D->pc) / 2.0 : 0.0;
  return VC_OK;
}

int vc_dist_n2ll(vc_handle* h, const double* theta, int32_t mu_mode, const double* mu_in,
                 double* n2ll, double* mu_out, double* logdet, double* quad) {
  VcDist* D = (VcDist*)h->dist;
  if (!D) return VC_ERR_INVALID;
  if (mu_mode != VC_MU_GLS && mu_mode != VC_MU_FIXED) return vc_fail(h, VC_ERR_INVALID, "unknown mu_mode");
  if (mu_mode == VC_MU_FIXED && !mu_in) return vc_fail(h, VC_ERR_INVALID, "mu_mode FIXED needs mu");
  VcTheta th;
  if (vc_make_theta(h, theta, &th) != VC_OK)
    return vc_fail(h, VC_ERR_INVALID, "theta does not define a covariance (range <= 0 or non-finite value)");
  try {
    VC_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->st_main;
    const int nt = D->nt, nb = D->nb, p = h->p;
    const int m = p + 1, npairs = m * (m + 1) / 2;
    if (mu_mode == VC_MU_FIXED)
      VC_CUDA_TRY(cudaMemcpyAsync(h->mu_in, mu_in, p * sizeof(double), cudaMemcpyHostToDevice, st));
    VC_CUDA_TRY(cudaEventRecord(h->ev[0], st));
    VC_CUDA_TRY(cudaMemsetAsync(h->cw.logdet_part, 0, (size_t)nt * sizeof(double), st));
    vc_launch_reset_info(h->cw.info, st, &h->launches);
    vc_launch_build_tiles(h->M, th, h->d, h->locs, h->w,
                          D->tiles_dev + D->build.off, D->build.n, D->pr, D->pc, st, &h->launches);
    if (D->my_pr == 0)
      vc_launch_border_dist(h->M.b, h->M.ldb, h->n_pad, p, h->x, h->y, D->col_tiles_dev, D->lct, st, &h->launches);
    VC_CUDA_TRY(cudaEventRecord(h->ev[1], st));

    VcGemmArgs base{};
    base.a = h->M.a; base.cb = h->M.cb; base.b = h->M.b; base.ldb = h->M.ldb; base.nt = nt;
    base.blk = nb; base.pr = D->pr; base.pc = D->pc;
    base.nbatch = 1;
    // Schedule of one outer block (lookahead; sm = main stream, sp = high-priority panel stream):
    //   sm: colD(s)  | colR(s) on num_sms - Rd ............ | rows(s+1) at FULL width | rest(s): part A on num_sms - Rn, part B on all
    //   sp:          '-> diag(s+1): L_DD chain, NCCL bcast -^             '-> NCCL panel broadcast(s+1), beside part A
    //  colD  update of the NEXT diagonal block only (its owner), so that its factorisation starts at once
    //  diag  owner factors L_DD (potrf / TRSM / inner updates on <= Rd SMs) and broadcasts it with the tile inverses
    //  rows  the ranks of the panel's process column solve their row tiles against L_DD (TRSM + inner updates)
    //        -- on EVERY SM, not on a reserved handful as in round 1 -- and pack their part of the panel
    //  rest  the trailing update; only the NCCL kernels of the panel broadcast need SMs beside it, and only for
    //        its first part (A); part B follows on all SMs and waits for nothing.
    // All NCCL calls are on sp, in the same order on every rank; panel buffers are double-buffered.
    const bool la = h->cw.lookahead;
    cudaStream_t sm = st, sp = la ? h->st_panel : st;
    if (la) {
      VC_CUDA_TRY(cudaEventRecord(h->cw.ev_update, sm));
      VC_CUDA_TRY(cudaStreamWaitEvent(sp, h->cw.ev_update, 0));
    }
    if (!D->ev_dd) {
      VC_CUDA_TRY(cudaEventCreateWithFlags(&D->ev_dd, cudaEventDisableTiming));
      VC_CUDA_TRY(cudaEventCreateWithFlags(&D->ev_rows, cudaEventDisableTiming));
    }
    static const int Rd = getenv("VC_DIST_RESERVE_DIAG") ? atoi(getenv("VC_DIST_RESERVE_DIAG")) : 8;
    const int Rn = vc_dist_nccl_ctas();
    static const double safety = getenv("VC_DIST_SAFETY") ? atof(getenv("VC_DIST_SAFETY")) : 1.5;
    double* dd_linv = D->dd + (size_t)nb * VC_TILE * ((size_t)nb * VC_TILE);
    // diagonal block of step i on stream q: factor (owner), broadcast L_DD + inverses
    auto panel_diag = [&](size_t i, cudaStream_t q) {
      const DistStep& s = D->steps[i];
      const int Kb = (s.oe - s.ob) * VC_TILE;
      if (D->rank == s.diag_owner) {
        for (int jt = s.ob; jt < s.oe; ++jt) {
          const int k = jt - s.ob;
          int64_t ldd;
          double* diag = vc_host_tile_ptr(h->M, loc_tile_h(jt, nb, D->pr), loc_tile_h(jt, nb, D->pc), &ldd);
          // potrf indexes its inverse slot by jt: bias the base so that slot jt is dd_linv[jt - ob]
          vc_launch_potrf(diag, ldd, jt, dd_linv - (int64_t)s.ob * VC_TILE * VC_TILE, h->cw.logdet_part,
                          h->cw.info, h->cw.legacy_potrf, q, &h->launches);
          VcGemmArgs g = base;
          g.linv = dd_linv + (size_t)k * VC_TILE * VC_TILE;
          g.jt = jt; g.K = VC_TILE;
          g.k0 = loc_tile_h(jt, nb, D->pc) * VC_TILE;
          g.tiles = D->tiles_dev + s.d_trsm[k].off; g.ntiles = s.d_trsm[k].n;
          vc_launch_trsm(g, g.ntiles, q, &h->launches);
          g.tiles = D->tiles_dev + s.d_inner[k].off; g.ntiles = s.d_inner[k].n;
          vc_launch_syrk(g, h->cw.num_sms, h->cw.legacy_syrk, q, &h->launches);
        }
        vc_launch_pack_tiles(base, D->tiles_dev + s.d_pack.off, s.d_pack.n,
                             loc_tile_h(s.ob, nb, D->pc) * VC_TILE, Kb, D->dd, s.ob, q, &h->launches);
      }
      if (D->world > 1)
        VC_NCCL_TRY(g_nccl.Broadcast(D->dd, D->dd, (size_t)nb * VC_TILE * ((size_t)nb * VC_TILE) +
                                                       (size_t)nb * VC_TILE * VC_TILE,
                                     ncclFloat64_, s.diag_owner, D->comm, q));
    };
    // row tiles of the panel of step i on stream q: solve against L_DD, pack
    auto panel_rows = [&](size_t i, cudaStream_t q) {
      const DistStep& s = D->steps[i];
      const int Kb = (s.oe - s.ob) * VC_TILE;
      if (D->my_pc != s.pcp || s.rows.n <= 0) return;
      for (int jt = s.ob; jt < s.oe; ++jt) {
        const int k = jt - s.ob;
        VcGemmArgs g = base;
        g.linv = dd_linv + (size_t)k * VC_TILE * VC_TILE;
        g.jt = jt; g.K = VC_TILE;
        g.k0 = loc_tile_h(jt, nb, D->pc) * VC_TILE;
        g.tiles = D->tiles_dev + s.rows.off; g.ntiles = s.rows.n;
        vc_launch_trsm(g, g.ntiles, q, &h->launches);
        g.tiles = D->tiles_dev + s.inner[k].off; g.ntiles = s.inner[k].n;
        g.src_b.pnl = D->dd; g.src_b.ld = VC_TILE; g.src_b.tile_stride = (int64_t)VC_TILE * Kb;
        g.src_b.tile0 = s.ob; g.src_b.col0 = k * VC_TILE;
        vc_launch_syrk(g, h->cw.num_sms, h->cw.legacy_syrk, q, &h->launches);
      }
      vc_launch_pack_tiles(base, D->tiles_dev + s.rows.off, s.rows.n,
                           loc_tile_h(s.ob, nb, D->pc) * VC_TILE, Kb, D->pnl[i & 1], s.oe, q, &h->launches, &s.map);
    };
    // every distribution block of the panel of step i goes from its owner to all ranks (stream q)
    auto panel_bcast = [&](size_t i, cudaStream_t q) {
      const DistStep& s = D->steps[i];
      double* pnl = D->pnl[i & 1];
      if (s.oe < nt && D->world > 1) {
        VC_NCCL_TRY(g_nccl.GroupStart());
        for (const BcastOp& op : s.bcasts)
          VC_NCCL_TRY(g_nccl.Broadcast(pnl + op.off, pnl + op.off, (size_t)op.count, ncclFloat64_, op.root,
                                       D->comm, q));
        VC_NCCL_TRY(g_nccl.GroupEnd());
      }
    };
    // whole panel of step i: diag on sp, rows on sm (they wait for L_DD), broadcast on sp (waits for the pack)
    auto panel = [&](size_t i) {
      panel_diag(i, sp);
      if (la) {
        VC_CUDA_TRY(cudaEventRecord(D->ev_dd, sp));
        VC_CUDA_TRY(cudaStreamWaitEvent(sm, D->ev_dd, 0));
      }
      panel_rows(i, sm);
      if (la) {
        VC_CUDA_TRY(cudaEventRecord(D->ev_rows, sm));
        VC_CUDA_TRY(cudaStreamWaitEvent(sp, D->ev_rows, 0));
      }
      panel_bcast(i, sp);
      if (la) VC_CUDA_TRY(cudaEventRecord(h->cw.ev_panel, sp));
    };
    auto update = [&](const VcGemmArgs& g0, const Span& sp_, int first, int count, int sms) {
      if (count <= 0) return;
      VcGemmArgs g = g0;
      g.tiles = D->tiles_dev + sp_.off + first; g.ntiles = count;
      vc_launch_syrk(g, sms, h->cw.legacy_syrk, sm, &h->launches);
    };
    const size_t nsteps = D->steps.size();
    const int nsm = h->cw.num_sms;
    panel(0);
    for (size_t i = 0; i < nsteps; ++i) {
      const DistStep& s = D->steps[i];
      if (s.oe >= nt) break;
      const int Kb = (s.oe - s.ob) * VC_TILE;
      if (la) VC_CUDA_TRY(cudaStreamWaitEvent(sm, h->cw.ev_panel, 0));       // panel i has arrived everywhere
      VcGemmArgs g = base;
      g.K = Kb;
      g.src_a.pnl = D->pnl[i & 1]; g.src_a.ld = VC_TILE; g.src_a.tile_stride = (int64_t)VC_TILE * Kb;
      g.src_a.tile0 = s.oe; g.src_a.col0 = 0;
      g.src_a.map = s.map;
      g.src_b = g.src_a;
      const bool more = i + 1 < nsteps;
      const bool ws = la && !h->cw.legacy_syrk;
      update(g, s.trailing_cold, 0, s.trailing_cold.n, nsm);                  // colD: the next diagonal block
      if (la) {
        VC_CUDA_TRY(cudaEventRecord(h->cw.ev_col, sm));
        VC_CUDA_TRY(cudaStreamWaitEvent(sp, h->cw.ev_col, 0));
      }
      if (more) {
        panel_diag(i + 1, sp);
        if (la) VC_CUDA_TRY(cudaEventRecord(D->ev_dd, sp));
      }
      update(g, s.trailing_col, 0, s.trailing_col.n, nsm - (ws && more ? std::min(Rd, nsm / 2) : 0));   // colR
      if (more) {
        if (la) VC_CUDA_TRY(cudaStreamWaitEvent(sm, D->ev_dd, 0));
        panel_rows(i + 1, sm);
        if (la) {
          VC_CUDA_TRY(cudaEventRecord(D->ev_rows, sm));
          VC_CUDA_TRY(cudaStreamWaitEvent(sp, D->ev_rows, 0));
        }
        panel_bcast(i + 1, sp);
        if (la) VC_CUDA_TRY(cudaEventRecord(h->cw.ev_panel, sp));
      }
      // rest(s): part A beside the NCCL broadcast of panel s+1 (sized from an estimate of its duration), part B on all SMs
      const int rest = s.trailing_rest.n;
      int nA = 0;
      const int reserve = std::max(0, std::min(Rn, nsm / 2));
      if (ws && more && D->world > 1 && reserve > 0) {
        const DistStep& nx = D->steps[i + 1];
        const double bytes = 8.0 * (double)(nt + D->nbt - nx.oe) * VC_TILE * (double)(nx.oe - nx.ob) * VC_TILE;
        const double t_bcast = 300.0 + bytes / 150e3 + 60.0 * (double)nx.bcasts.size();   // us: ~150 GB/s + per-broadcast latency
        const double t_tile = 17.3 * (Kb / VC_TILE);
        nA = (int)std::min<int64_t>(rest, (int64_t)(safety * t_bcast / t_tile + 1.0) * (nsm - reserve));
      }
      update(g, s.trailing_rest, 0, nA, nsm - reserve);
      update(g, s.trailing_rest, nA, rest - nA, nsm);
    }
    if (la) VC_CUDA_TRY(cudaStreamWaitEvent(sm, h->cw.ev_panel, 0));
    VC_CUDA_TRY(cudaEventRecord(h->ev[2], st));

    // ---- finalize: local partials, three small all-reduces ----
    double* red_gram = D->red;
    double* red_logdet = D->red + 2 * npairs;
    double* red_quad = red_logdet + 2;
    if (D->my_pr == 0 && D->lct > 0)
      vc_launch_gram_local(h->M.b, h->M.ldb, (int64_t)D->lct * VC_TILE, p, h->fw, red_gram, st, &h->launches);
    else
      VC_CUDA_TRY(cudaMemsetAsync(red_gram, 0, 2 * npairs * sizeof(double), st));
    vc_launch_sum(h->cw.logdet_part, nt, red_logdet, st, &h->launches);
    if (D->world > 1) {
      VC_NCCL_TRY(g_nccl.AllReduce(D->red, D->red, 2 * npairs + 1, ncclFloat64_, ncclSum_, D->comm, st));
      VC_NCCL_TRY(g_nccl.AllReduce(h->cw.info, D->info_red, 1, ncclInt32_, ncclMin_, D->comm, st));
    } else {
      VC_CUDA_TRY(cudaMemcpyAsync(D->info_red, h->cw.info, sizeof(int), cudaMemcpyDeviceToDevice, st));
    }
    vc_launch_solve_reduced(red_gram, p, mu_mode, h->mu_in, red_logdet, D->info_red, h->results, h->gram_dev,
                            st, &h->launches);
    if (D->my_pr == 0 && D->lct > 0)
      vc_launch_quad_local(h->M.b, h->M.ldb, (int64_t)D->lct * VC_TILE, p, h->results, h->fw, red_quad, st,
                           &h->launches);
    else
      VC_CUDA_TRY(cudaMemsetAsync(red_quad, 0, 2 * sizeof(double), st));
    if (D->world > 1)
      VC_NCCL_TRY(g_nccl.AllReduce(red_quad, red_quad, 2, ncclFloat64_, ncclSum_, D->comm, st));
    vc_launch_assemble_reduced(red_quad, h->n, h->results, st, &h->launches);
    VC_CUDA_TRY(cudaEventRecord(h->ev[3], st));
    VC_CUDA_TRY(cudaMemcpyAsync(h->results_host, h->results, VC_RES_STRIDE * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    VC_CUDA_TRY(cudaStreamSynchronize(st));
    VC_CUDA_TRY(cudaGetLastError());
    const double* r = h->results_host;
    float b = 0, c = 0, f = 0;
    cudaEventElapsedTime(&b, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&c, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&f, h->ev[2], h->ev[3]);
    h->last_timing = {b, c, f, b + c + f};
    *n2ll = r[VC_RES_N2LL];
    if (mu_out) for (int i = 0; i < p; ++i) mu_out[i] = r[VC_RES_MU + i];
    if (logdet) *logdet = r[VC_RES_LOGDET];
    if (quad) *quad = r[VC_RES_QUAD];
    const int info = (int)r[VC_RES_INFO];
    h->last_info = info;
    if (info != 0) {
      char buf[128];
      snprintf(buf, sizeof(buf), "the leading minor of order %d is not positive", info);
      h->err = buf;
      return VC_ERR_NOT_PD;
    }
  } catch (const VcError& e) {
    return vc_fail(h, e.code, e.msg);
  }
  return VC_OK;
}
