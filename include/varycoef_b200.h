/*
 * varycoef_b200.h -- C ABI of the B200-native maximum-likelihood objective of
 * GP-based spatially varying coefficient (SVC) models.
 *
 * This is the drop-in boundary for varycoef's hot path
 *     n2LL -> Sigma_y -> chol -> GLS_chol -> -2 logL
 * (reference: R-pkg/R/objective_functions.R:4-62, R-pkg/R/utils.R:77-101,171-233,508-552).
 * The reference has no FFI of its own (pure R); the boundary it exposes is the
 * closure `fn(x, ...) -> numeric(1)` handed to stats::optim
 * (R-pkg/R/SVC_mle.R:145-163) and exported through
 * SVC_mle_control(extract_fun = TRUE) (R-pkg/R/SVC_mle.R:111-132).  Every entry
 * point below names the reference code it replaces; INTEGRATION.md shows the
 * `.Call` stubs a varycoef maintainer would add.
 *
 * Conventions: all matrices are R's column-major `double`; the caller owns every
 * host buffer; the library copies inputs to the device in vc_create and never
 * retains host pointers; a handle is opaque, not thread-safe, and is freed with
 * vc_destroy.  Every function returns a status and never throws / exits.
 *
 * Parameter vector layout (R-pkg/R/utils.R:111-115):
 *     theta = (range_1, var_1, ..., range_q, var_q, nugget_var)      length 2q+1
 */
#ifndef VARYCOEF_B200_H
#define VARYCOEF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC_ABI_VERSION 1

#if defined(__GNUC__)
#define VC_API __attribute__((visibility("default")))
#else
#define VC_API
#endif

/* status codes */
#define VC_OK 0
#define VC_ERR_NOT_PD 1   /* base::chol: "the leading minor of order k is not positive"; k via vc_last_info */
#define VC_ERR_INVALID 2  /* bad argument, or theta fails check_cov_lower (R-pkg/R/utils.R:291-297) */
#define VC_ERR_CUDA 3     /* CUDA / NCCL failure, message via vc_last_error */
#define VC_ERR_OOM 4      /* device or host allocation failed */

/* cov.name of SVC_mle_control (R-pkg/R/SVC_mle.R:428), mapped by MLE.cov.func (R-pkg/R/utils.R:18-29) */
enum { VC_COV_EXP = 0, VC_COV_MAT32 = 1, VC_COV_MAT52 = 2, VC_COV_SPH = 3, VC_COV_WEND1 = 4, VC_COV_WEND2 = 5 };
/* control$dist$method, passed to stats::dist / spam::nearest.dist (R-pkg/R/utils.R:514-538) */
enum { VC_DIST_EUCLIDEAN = 0, VC_DIST_MAXIMUM = 1, VC_DIST_MANHATTAN = 2, VC_DIST_MINKOWSKI = 3 };
/* get_taper (R-pkg/R/utils.R:545-552): exp, mat32 -> wend1; mat52 -> wend2 */
enum { VC_TAPER_NONE = 0, VC_TAPER_WEND1 = 1, VC_TAPER_WEND2 = 2 };
/* how mu enters n2LL (R-pkg/R/objective_functions.R:31-43) */
enum { VC_MU_GLS = 0,   /* profile && is.null(mean.est): GLS_chol */
       VC_MU_FIXED = 1  /* mean.est given, or non-profile: mu = x[2q+1 + 1:p] */ };

typedef struct vc_handle vc_handle;

typedef struct vc_config {
  int32_t cov_id;        /* VC_COV_*  */
  int32_t dist_method;   /* VC_DIST_* */
  double  minkowski_p;   /* only for VC_DIST_MINKOWSKI */
  int32_t taper_id;      /* VC_TAPER_*; control$tapering != NULL  */
  int32_t device;        /* CUDA device ordinal */
  double  taper_range;   /* control$tapering (delta) */
  int32_t nb_outer;      /* outer Cholesky block (multiple of 128; 0 = auto: 128 / 256 / 512 / 1024 by n) */
  int32_t flags;         /* VC_FLAG_* */
} vc_config;

#define VC_FLAG_NO_LOOKAHEAD 1 /* serialise panel factorisation and trailing update (debug) */
#define VC_FLAG_LEGACY_POTRF 4 /* unblocked diagonal-tile factorisation kernel (A/B) */
#define VC_FLAG_LEGACY_SYRK 2  /* one-CTA-per-tile cp.async update kernel instead of the persistent one (A/B) */

typedef struct vc_timing {   /* device times of the LAST evaluation, milliseconds (CUDA events) */
  float build_ms;            /* fused distance + covariance build (Sigma_y)         */
  float chol_ms;             /* bordered blocked Cholesky (chol + the L^-1 [X y])   */
  float finalize_ms;         /* GLS normal equations, quadratic form, log-det       */
  float total_ms;
} vc_timing;

/* fills cfg with the defaults of SVC_mle_control(): exp covariance, Euclidean distance, no taper */
VC_API int vc_config_default(vc_config* cfg);

/*
 * Replaces the per-fit set-up of MLE_computation (R-pkg/R/SVC_mle.R:33-36 own_dist,
 * :58-61 cov.func closure, :65-72 outer.W): nothing n x n is precomputed, the O(n)
 * inputs are copied to the device.   locs: n x d, X: n x p, W: n x q (column-major), y: n.
 */
VC_API int vc_create(vc_handle** h, const vc_config* cfg, int64_t n, int32_t d, int32_t p, int32_t q,
              const double* locs, const double* X, const double* W, const double* y);

/*
 * The same objective on SEVERAL GPUs of one process: one replica of the O(n) inputs per device;
 * vc_n2ll_batch spreads the distinct thetas of a batch over the devices (no collective; results are gathered on
 * the host).  This is the GPU analogue of the reference's optimParallel fan-out over PSOCK workers
 * (R-pkg/R/SVC_mle.R:164-200, control$parallel): optim's finite-difference and Hessian stencils are evaluated by
 * all devices at once.  Every other entry point acts on the device that ran the last evaluation.
 * cfg->device is ignored; devices[] may name a device more than once.
 */
VC_API int vc_create_multi(vc_handle** h, const vc_config* cfg, int64_t n, int32_t d, int32_t p, int32_t q,
                    const double* locs, const double* X, const double* W, const double* y,
                    const int32_t* devices, int32_t ndev);
VC_API int32_t vc_device_count(vc_handle* h);       /* devices behind the handle (1 unless vc_create_multi) */

/* re-upload the O(n) inputs of an existing handle (same shapes); any pointer may be NULL = keep */
VC_API int vc_set_data(vc_handle* h, const double* locs, const double* X, const double* W, const double* y);

VC_API void vc_destroy(vc_handle* h);

/*
 * One objective evaluation: n2LL (R-pkg/R/objective_functions.R:4-62) WITHOUT the
 * pc_penalty term (host scalar, stays in R: objective_functions.R:110-123) and WITH
 * n*log(2 pi).   theta: 2q+1.   mu_mode VC_MU_GLS: mu_in ignored, GLS estimate
 * (GLS_chol, R-pkg/R/utils.R:93-101) returned in mu_out; VC_MU_FIXED: mu_in (p) is used.
 * Optional outputs (may be NULL): mu_out (p), logdet = log|Sigma_y|, quad = r' Sigma^-1 r.
 */
VC_API int vc_n2ll(vc_handle* h, const double* theta, int32_t mu_mode, const double* mu_in,
            double* n2ll, double* mu_out, double* logdet, double* quad);

/*
 * m evaluations queued back to back with a single synchronisation -- the stencil
 * optim's numeric gradient / optimhess / optimParallel evaluate
 * (R-pkg/R/SVC_mle.R:145-199).  thetas: (2q+1) x m column-major; mus_in: p x m or NULL;
 * out_n2ll: m; out_mu: p x m or NULL; out_status: m (VC_OK / VC_ERR_NOT_PD / VC_ERR_INVALID)
 * or NULL.  Returns the first non-OK status (all m are still attempted).
 */
VC_API int vc_n2ll_batch(vc_handle* h, int32_t m, const double* thetas, int32_t mu_mode,
                  const double* mus_in, double* out_n2ll, double* out_mu, int32_t* out_status);

/* Sigma_y(theta) (R-pkg/R/utils.R:171-233) as a full symmetric n x n column-major matrix (parity/debug) */
VC_API int vc_sigma(vc_handle* h, const double* theta, double* out);
/* the same matrix written to DEVICE memory of the handle's GPU (n x n column-major, ld = n): lets a caller
 * hand Sigma_y to another device library (cuSOLVER, as the parity tests do at n = 50 000) without a host trip */
VC_API int vc_sigma_device(vc_handle* h, const double* theta, double* out_dev);

/*
 * Cholesky factor of the last successful evaluation, as base::chol returns it:
 * upper triangular R (n x n column-major, zeros below the diagonal), R'R = Sigma_y
 * (R-pkg/R/objective_functions.R:26).
 */
VC_API int vc_get_chol(vc_handle* h, double* R_out);

/* Z = R^-T [X y] of the last evaluation (n x (p+1) column-major) -- solve(t(R), .) of R-pkg/R/utils.R:94,100 */
VC_API int vc_get_whitened(vc_handle* h, double* Z_out);

/*
 * GLS_chol.matrix (R-pkg/R/utils.R:93-101, exported NAMESPACE:24) for a caller-supplied
 * upper Cholesky factor R (n x n column-major): mu = (X' S^-1 X)^-1 X' S^-1 y.
 * Stand-alone (no handle): runs on `device`.
 */
VC_API int vc_gls_chol(int32_t device, int64_t n, int32_t p, const double* R, const double* X,
                const double* y, double* mu_out);

/*
 * Next-tier rows, all reusing the resident factor of theta:
 *  vc_post    : prep_par_output / eff_dof (R-pkg/R/utils.R:423-473, R-pkg/R/eff_dof.R:7-21):
 *               mu_gls (p), Sigma_FE = (X' S^-1 X)^-1 (p x p), edof.  Any output may be NULL.
 *  vc_predict : predict.SVC_mle (R-pkg/R/predict-SVC_mle.R:80-267): EBLUP of the q SVC
 *               deviations at n_new locations (eff: n_new x q), optional y_pred (needs newX, newW)
 *               and y_var (compute.y.var = TRUE).  newlocs: n_new x d.
 */
VC_API int vc_post(vc_handle* h, const double* theta, double* mu_gls, double* sigma_fe, double* edof);
VC_API int vc_predict(vc_handle* h, const double* theta, const double* mu, int64_t n_new,
               const double* newlocs, const double* newX, const double* newW,
               double* eff, double* y_pred, double* y_var);

/*
 * med_dist of MLE_computation (R-pkg/R/SVC_mle.R:39-45, dense branch): median(as.numeric(d)) over the full
 * n x n distance matrix (zero diagonal and both triangles included), from which the default init / bounds of
 * the optimiser are derived (R-pkg/R/utils.R:336-396).  Exact (radix select on recomputed tile-local
 * distances), nothing n x n is held.  Stand-alone: no handle.  locs: n x d column-major.
 */
VC_API int vc_median_dist(int32_t device, int64_t n, int32_t d, const double* locs, int32_t dist_method,
                          double minkowski_p, double* med_out, int32_t* launches_out);

/* diagnostics */
VC_API int vc_timings(vc_handle* h, vc_timing* out);
VC_API int vc_last_info(vc_handle* h);              /* order of the failing leading minor, 0 if none */
VC_API const char* vc_last_error(vc_handle* h);     /* h may be NULL: error of the last failed vc_create */
VC_API int64_t vc_padded_n(vc_handle* h);           /* n rounded up to the 128-row tile */
VC_API int64_t vc_launch_count(vc_handle* h);       /* kernels launched by this handle so far */
VC_API int vc_abi_version(void);

/* measured FP64 tensor (DMMA) and HBM-write peaks of `device`, used as roofline denominators */
VC_API int vc_probe_dmma_peak(int32_t device, int32_t iters, double* tflops);
VC_API int vc_probe_hbm_write(int32_t device, int64_t bytes, int32_t iters, double* gbs);

/*
 * Multi-GPU, one process per GPU: Sigma_y built and factored 2D
 * block-cyclic on a pr x pc process grid, panels broadcast with NCCL.
 * vc_nccl_unique_id: rank 0 fills 128 bytes that the launcher ships to all ranks
 * (torch.distributed / MPI / files); every rank then calls vc_create_dist with the
 * same id, grid and inputs.  vc_n2ll & co. are collective on such a handle.
 */
VC_API int vc_nccl_unique_id(char id_out[128]);
VC_API int vc_create_dist(vc_handle** h, const vc_config* cfg, int64_t n, int32_t d, int32_t p, int32_t q,
                   const double* locs, const double* X, const double* W, const double* y,
                   int32_t rank, int32_t world, int32_t pr, int32_t pc, const char nccl_id[128]);

/* host-only: order of the row tiles >= oe in the gathered panel buffer of the sharded path (owner-major: process row 0,
 * its border tiles, process row 1, ...): first position / tile count per process row, position of every row tile
 * oe .. nt + nbt - 1.  Every process row's part is one contiguous range = one broadcast.  Returns the tile count. */
VC_API int vc_bc_panel_layout(int32_t nt, int32_t nbt, int32_t nb, int32_t pr, int32_t oe, int32_t* chunk_off,
                              int32_t* chunk_cnt, int32_t* pos);
/* NVLink bytes this rank receives per evaluation (panel + L_DD broadcasts on the world communicator), next to the
 * n^2 8 (1/pr + 1/pc) / 2 a row / column-communicator scheme would need */
VC_API int vc_dist_traffic(vc_handle* h, double* rx_bytes, double* rowcol_bound_bytes);

/* host-only description of the lower block-column storage of an n x n matrix (no GPU needed; unit tests and sizing):
 * block column c starts `off[c]` doubles into the slot, has leading dimension ld[c] and stores the row tiles from r0[c]
 * on.  Returns the number of block columns (arrays are filled up to cap); *doubles_out = doubles per evaluation slot. */
VC_API int vc_layout_describe(int64_t n, int32_t nb_outer, int64_t* doubles_out, int64_t* off, int32_t* ld,
                              int32_t* r0, int32_t cap);

/* host-only plan of a batch of m evaluations under S evaluation slots (no GPU needed; unit tests): evaluations that
 * repeat theta share a slot, the last valid evaluation lands in slot 0 of the last group.  valid may be NULL.
 * Returns the number of distinct thetas; group_of / slot_of / rec_of (m each) are -1 for invalid evaluations. */
VC_API int vc_plan_batch(int32_t m, int32_t nth, const double* thetas, const int32_t* valid, int32_t S,
                         int32_t* group_of, int32_t* slot_of, int32_t* rec_of);

/* host-only block-cyclic map (no GPU needed): owner rank and local tile index of 512-block (bi, bj) */
VC_API int vc_bc_owner(int32_t bi, int32_t bj, int32_t pr, int32_t pc);
VC_API int vc_bc_local(int32_t b, int32_t nprocs_dim);
VC_API int vc_bc_count(int32_t nblocks, int32_t coord, int32_t nprocs_dim);
/* trailing-update tiles (packed I | J << 16, global 128-tile indices) owned by `rank` after the outer block
 * ending at tile `oe`: J in [oe, nt), I in [J, nt + nbt); returns the count, fills out[0..cap) if non-NULL */
VC_API int vc_bc_trailing_tiles(int32_t nt, int32_t nbt, int32_t nb, int32_t pr, int32_t pc, int32_t rank,
                         int32_t oe, int32_t* out, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* VARYCOEF_B200_H */
