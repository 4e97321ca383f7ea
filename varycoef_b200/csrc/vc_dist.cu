// vc_dist.cu -- K4 (multi-GPU 2D block-cyclic build + Cholesky with NCCL panel broadcasts).
// Placeholder while the single-GPU path is brought up; the host-only block-cyclic maps are final.
#include "vc_handle.cuh"

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

extern "C" int vc_nccl_unique_id(char id_out[128]) { (void)id_out; return VC_ERR_INVALID; }
extern "C" int vc_create_dist(vc_handle** h, const vc_config*, int64_t, int32_t, int32_t, int32_t,
                              const double*, const double*, const double*, const double*, int32_t,
                              int32_t, int32_t, int32_t, const char*) {
  if (h) *h = nullptr;
  return vc_fail(nullptr, VC_ERR_INVALID, "vc_create_dist: not built yet");
}
int vc_dist_n2ll(vc_handle* h, const double*, int32_t, const double*, double*, double*, double*, double*) {
  return vc_fail(h, VC_ERR_INVALID, "distributed path not built yet");
}
void vc_dist_destroy(vc_handle*) {}
