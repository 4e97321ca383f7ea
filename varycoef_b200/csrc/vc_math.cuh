// vc_math.cuh -- element-wise fp64 arithmetic of the covariance build (host + device).
//
// Covariance shapes follow spam::cov.exp / cov.mat / cov.sph / cov.wend1 / cov.wend2 as
// dispatched by MLE.cov.func (reference R-pkg/R/utils.R:2-29); t = h / range.  All shapes
// equal 1 at t = 0, so spam's `t < eps -> sill + nugget` branch needs no special case.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define VC_HD __host__ __device__ __forceinline__
#else
#define VC_HD inline
#endif

#define VC_EXP_TABLE_BITS 6
#define VC_EXP_TABLE_SIZE 64

// 2^(j/64), j = 0..63 (filled on the host with exp2(), copied to shared memory by kernels)
static inline void vc_fill_exp2_table(double* t) {
  for (int j = 0; j < VC_EXP_TABLE_SIZE; ++j) t[j] = exp2((double)j / VC_EXP_TABLE_SIZE);
}

VC_HD double vc_hiloint2double(int32_t hi, int32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}
VC_HD int32_t vc_double2loint(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return (int32_t)(uint32_t)u;
#endif
}
VC_HD int32_t vc_double2hiint(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return (int32_t)(uint32_t)(u >> 32);
#endif
}

// exp(-t) for t >= 0, ~1 ulp.  Table-driven: -t = (64 m + j) ln2/64 + r, |r| <= ln2/128,
// exp(-t) = 2^m * 2^(j/64) * (1 + r + r^2/2 + ... + r^6/720).  About 12 fp64 issue slots
// against ~30 for the libdevice exp(): the build kernel is fp64-ALU-bound, not HBM-bound.
VC_HD double vc_exp_neg(double t, const double* __restrict__ tab) {
  // branch-free: evaluate on min(t, 708) and select at the end (exp(-708) = 3e-308: flush below)
  const double x = -fmin(t, 708.0);
  const double kMagic = 6755399441055744.0;          // 2^52 + 2^51
  const double kInv = 92.332482616893656;            // 64 / ln 2
  const double kLn2Hi = 6.93147180369123816490e-01 / 64.0;
  const double kLn2Lo = 1.90821492927058770002e-10 / 64.0;
  double kd = fma(x, kInv, kMagic);
  const int32_t ki = vc_double2loint(kd);
  kd -= kMagic;
  double r = fma(kd, -kLn2Hi, x);
  r = fma(kd, -kLn2Lo, r);
  double p = 1.0 / 720.0;
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = p * r;                                          // e^r - 1
  const double tj = tab[ki & (VC_EXP_TABLE_SIZE - 1)];
  double v = fma(tj, p, tj);
  const int32_t m = ki >> VC_EXP_TABLE_BITS;           // arithmetic shift, m in [-1022, 0]
  v = vc_hiloint2double(vc_double2hiint(v) + (m << 20), vc_double2loint(v));
  return (t < 708.0) ? v : ((t != t) ? t : 0.0);
}

// sqrt(s) for s >= 0, <= 1 ulp, branch-free on the device (reciprocal square root seed + two Newton
// steps + one correction); the libdevice sqrt carries a slow-path call that breaks up the inner loop.
// (One Newton step would do arithmetically; measured, it does not shorten the kernel -- 1510 vs 1524 GB/s --
// so the arithmetic validated in round 1 is kept.)
VC_HD double vc_sqrt_pos(double s) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  const double hs = 0.5 * s;
  y = y * fma(-hs * y, y, 1.5);                        // Newton on 1/sqrt(s)
  y = y * fma(-hs * y, y, 1.5);
  double h = s * y;
  h = fma(fma(-h, h, s), 0.5 * y, h);                  // h += (s - h^2) / (2 sqrt(s))
  return (s > 2.3e-308) ? h : 0.0;
#else
  return sqrt(s);
#endif
}

// shape(t) = cov(h; range, sill = 1), t = h / range >= 0
template <int COV>
VC_HD double vc_cov_shape(double t, const double* __restrict__ tab) {
  if (COV == 0) return vc_exp_neg(t, tab);                                  // cov.exp
  if (COV == 1) return (1.0 + t) * vc_exp_neg(t, tab);                      // cov.mat, nu = 3/2
  if (COV == 2) return fma(t, fma(t, 1.0 / 3.0, 1.0), 1.0) * vc_exp_neg(t, tab);  // nu = 5/2
  if (COV == 3) return t < 1.0 ? fma(0.5 * t * t, t, fma(-1.5, t, 1.0)) : 0.0;    // cov.sph
  if (COV == 4) {                                                           // cov.wend1
    if (!(t < 1.0)) return 0.0;
    double u = 1.0 - t, u2 = u * u;
    return u2 * u2 * fma(4.0, t, 1.0);
  }
  {                                                                         // cov.wend2
    if (!(t < 1.0)) return 0.0;
    double u = 1.0 - t, u2 = u * u;
    return u2 * u2 * u2 * fma(t, fma(35.0, t, 18.0), 3.0) * (1.0 / 3.0);
  }
}

VC_HD double vc_cov_shape_rt(int cov, double t, const double* __restrict__ tab) {
  switch (cov) {
    case 0: return vc_cov_shape<0>(t, tab);
    case 1: return vc_cov_shape<1>(t, tab);
    case 2: return vc_cov_shape<2>(t, tab);
    case 3: return vc_cov_shape<3>(t, tab);
    case 4: return vc_cov_shape<4>(t, tab);
    default: return vc_cov_shape<5>(t, tab);
  }
}

// taper(h) of get_taper (R-pkg/R/utils.R:545-552): cov.wend{1,2}(h, c(delta, 1, 0))
VC_HD double vc_taper(int taper_id, double h, double inv_delta) {
  if (taper_id == 0) return 1.0;
  const double t = h * inv_delta;
  return taper_id == 1 ? vc_cov_shape<4>(t, nullptr) : vc_cov_shape<5>(t, nullptr);
}

// accumulate one coordinate into the running distance (stats::dist C code: direct differences)
VC_HD double vc_dist_accum(int method, double acc, double dev, double mink_p) {
  dev = fabs(dev);
  switch (method) {
    case 0: return fma(dev, dev, acc);
    case 1: return fmax(acc, dev);
    case 2: return acc + dev;
    default: return acc + pow(dev, mink_p);
  }
}
VC_HD double vc_dist_finish(int method, double acc, double mink_p) {
  switch (method) {
    case 0: return sqrt(acc);
    case 3: return pow(acc, 1.0 / mink_p);
    default: return acc;
  }
}
