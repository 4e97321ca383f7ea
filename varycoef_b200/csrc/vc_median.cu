// vc_median.cu -- med_dist of MLE_computation on the GPU.
//
// Replaces  med_dist <- median(as.numeric(d))  (reference R-pkg/R/SVC_mle.R:39-45, dense branch): the median
// of ALL n^2 entries of the distance matrix -- n zeros on the diagonal and every pair distance twice -- from
// which the optimiser's default init / bounds are derived (R-pkg/R/utils.R:336-396).  The reference needs the
// materialised n x n matrix for it (20 GB at n = 50 000); here the distances are recomputed tile by tile and the
// order statistic is found EXACTLY by a radix select on the bit patterns of the (non-negative) doubles:
// six passes of 11/11/11/11/10/10 bits, each a histogram of the pairs that match the prefix found so far
// (shared-memory histogram per CTA, flushed with 64-bit atomics).  Distances are formed exactly as
// stats::dist does (direct differences, multiply and add rounded separately, IEEE sqrt), so the result is
// bit-identical to the host computation.
#include "vc_handle.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {

constexpr int MT = 128;          // pairs tile edge
constexpr int MTHREADS = 256;
constexpr int MAXBINS = 2048;

struct MedArgs {
  const double* locs;   // n x d column-major
  int64_t n;
  int d;
  int method;
  double mink_p;
  int shift;            // bit position of the digit of this pass
  int bits;             // digit width
  unsigned long long prefix;   // bits above shift + bits must equal this (already shifted down)
  int has_prefix;
  unsigned long long* hist;    // 2^bits global bins
};

__device__ __forceinline__ double med_dist_pair(const MedArgs& p, const double* xi, const double* sx, int jj) {
  double acc = 0.0;
  for (int c = 0; c < p.d; ++c) {
    const double dev = fabs(xi[c] - sx[c * MT + jj]);
    switch (p.method) {
      case VC_DIST_EUCLIDEAN: acc = __dadd_rn(acc, __dmul_rn(dev, dev)); break;   // no FMA: as the C code of dist()
      case VC_DIST_MAXIMUM: acc = fmax(acc, dev); break;
      case VC_DIST_MANHATTAN: acc = __dadd_rn(acc, dev); break;
      default: acc = __dadd_rn(acc, pow(dev, p.mink_p)); break;
    }
  }
  if (p.method == VC_DIST_EUCLIDEAN) return __dsqrt_rn(acc);
  if (p.method == VC_DIST_MINKOWSKI) return pow(acc, 1.0 / p.mink_p);
  return acc;
}

// one CTA per lower tile (I >= J) of the pair matrix; strictly-lower pairs only
__global__ void __launch_bounds__(MTHREADS) k_median_hist(MedArgs p) {
  __shared__ unsigned int sh[MAXBINS];
  __shared__ double sx[VC_MAXD * MT];
  const int id = blockIdx.x;
  int I = (int)((sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
  while ((I + 1) * (I + 2) / 2 <= id) ++I;
  while (I * (I + 1) / 2 > id) --I;
  const int J = id - I * (I + 1) / 2;
  const int tid = threadIdx.x;
  const int nbins = 1 << p.bits;
  for (int i = tid; i < nbins; i += MTHREADS) sh[i] = 0u;
  for (int idx = tid; idx < p.d * MT; idx += MTHREADS) {
    const int c = idx / MT, r = idx % MT;
    const int64_t j = (int64_t)J * MT + r;
    sx[idx] = j < p.n ? p.locs[j + (int64_t)c * p.n] : 0.0;
  }
  __syncthreads();
  const int r = tid & (MT - 1), half = tid >> 7;
  const int64_t i = (int64_t)I * MT + r;
  double xi[VC_MAXD];
  for (int c = 0; c < p.d; ++c) xi[c] = i < p.n ? p.locs[i + (int64_t)c * p.n] : 0.0;
  if (i < p.n) {
    for (int jj = half * 64; jj < half * 64 + 64; ++jj) {
      const int64_t j = (int64_t)J * MT + jj;
      if (j >= i || j >= p.n) break;                 // strictly lower pairs (j < i); columns ascend
      const unsigned long long key = (unsigned long long)__double_as_longlong(med_dist_pair(p, xi, sx, jj));
      if (p.has_prefix && (key >> (p.shift + p.bits)) != p.prefix) continue;
      atomicAdd(&sh[(unsigned)((key >> p.shift) & (unsigned long long)(nbins - 1))], 1u);
    }
  }
  __syncthreads();
  for (int b = tid; b < nbins; b += MTHREADS)
    if (sh[b]) atomicAdd(&p.hist[b], (unsigned long long)sh[b]);
}

// k-th smallest (0-based) of the n (n - 1) / 2 pair distances
double select_pair_distance(MedArgs a, unsigned long long k, cudaStream_t st) {
  static const int widths[6] = {11, 11, 11, 11, 10, 10};     // 64 bits, most significant digit first
  const int64_t nt = (a.n + MT - 1) / MT;
  const unsigned grid = (unsigned)(nt * (nt + 1) / 2);
  std::vector<unsigned long long> h(MAXBINS);
  unsigned long long prefix = 0;
  int shift = 64;
  for (int pass = 0; pass < 6; ++pass) {
    const int bits = widths[pass];
    shift -= bits;
    a.shift = shift; a.bits = bits; a.prefix = prefix; a.has_prefix = pass > 0;
    VC_CUDA_TRY(cudaMemsetAsync(a.hist, 0, MAXBINS * sizeof(unsigned long long), st));
    k_median_hist<<<grid, MTHREADS, 0, st>>>(a);
    VC_CUDA_TRY(cudaMemcpyAsync(h.data(), a.hist, (size_t)(1 << bits) * sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, st));
    VC_CUDA_TRY(cudaStreamSynchronize(st));
    unsigned long long cum = 0;
    int bin = 0;
    for (; bin < (1 << bits); ++bin) {
      if (cum + h[bin] > k) break;
      cum += h[bin];
    }
    if (bin == (1 << bits)) throw VcError(VC_ERR_CUDA, "median select: rank beyond the histogram (non-finite coordinates?)");
    k -= cum;
    prefix = (prefix << bits) | (unsigned long long)bin;
  }
  double out;
  static_assert(sizeof(out) == sizeof(prefix), "64-bit double");
  memcpy(&out, &prefix, sizeof(out));
  return out;
}

}  // namespace

extern "C" int vc_median_dist(int32_t device, int64_t n, int32_t d, const double* locs, int32_t dist_method,
                              double minkowski_p, double* med_out, int32_t* launches_out) {
  if (!locs || !med_out || n < 1 || d < 1 || d > VC_MAXD) return VC_ERR_INVALID;
  if (dist_method < 0 || dist_method > VC_DIST_MINKOWSKI) return VC_ERR_INVALID;
  if (dist_method == VC_DIST_MINKOWSKI && !(minkowski_p > 0.0)) return VC_ERR_INVALID;
  double* dl = nullptr;
  unsigned long long* hist = nullptr;
  cudaStream_t st = nullptr;
  int rc = VC_OK;
  try {
    VC_CUDA_TRY(cudaSetDevice(device));
    VC_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    VC_CUDA_TRY(cudaMalloc(&dl, (size_t)n * d * sizeof(double)));
    VC_CUDA_TRY(cudaMalloc(&hist, MAXBINS * sizeof(unsigned long long)));
    VC_CUDA_TRY(cudaMemcpyAsync(dl, locs, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, st));
    MedArgs a{};
    a.locs = dl; a.n = n; a.d = d; a.method = dist_method; a.mink_p = minkowski_p; a.hist = hist;
    // order statistics (0-based) of the full multiset = n zeros followed by every pair distance twice
    const unsigned long long N = (unsigned long long)n * (unsigned long long)n;
    unsigned long long ranks[2];
    int nr;
    if (N % 2 == 1) { ranks[0] = N / 2; nr = 1; } else { ranks[0] = N / 2 - 1; ranks[1] = N / 2; nr = 2; }
    double vals[2] = {0.0, 0.0};
    int launches = 0;
    for (int r = 0; r < nr; ++r) {
      if (ranks[r] < (unsigned long long)n) { vals[r] = 0.0; continue; }
      const unsigned long long k = (ranks[r] - (unsigned long long)n) / 2;
      if (r == 1 && ranks[0] >= (unsigned long long)n && (ranks[0] - (unsigned long long)n) / 2 == k) { vals[1] = vals[0]; continue; }
      vals[r] = select_pair_distance(a, k, st);
      launches += 6;
    }
    *med_out = nr == 1 ? vals[0] : (vals[0] + vals[1]) * 0.5;      // median() averages the two middle values
    if (launches_out) *launches_out = launches;
  } catch (const VcError& e) {
    rc = vc_fail(nullptr, e.code, e.msg);
  }
  cudaFree(dl); cudaFree(hist);
  if (st) cudaStreamDestroy(st);
  return rc;
}
