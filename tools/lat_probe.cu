// Dependent-issue latencies that bound the diagonal-tile kernel (development aid):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/lat_probe.cu -o tools/bin/lat_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, long long* cyc, double seed, int iters) {
  double x = seed + threadIdx.x * 1e-9, y = 1.0 - 1e-9 * threadIdx.x, c1 = 0.0;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) x = fma(x, y, 1e-3);                                        // DFMA chain
      if (MODE == 1) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);            // 64-bit shuffle chain
      if (MODE == 2) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }  // MUFU + DADD
      if (MODE == 3) dmma884(x, c1, y, y);                                        // DMMA chain through C
      if (MODE == 4) x = x * y;                                                   // DMUL chain
      if (MODE == 5) x = sqrt(x) + 1.0;                                           // IEEE sqrt + DADD
      if (MODE == 6) x = 1.0 / x + 1.0;                                           // IEEE div + DADD
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + c1;
}
template <int MODE> void run(const char* name, int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  k<MODE><<<1, 32 * warps>>>(out, cyc, 1.000001, iters);
  k<MODE><<<1, 32 * warps>>>(out, cyc, 1.000001, iters);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps=%d  %.1f cycles / dependent op\n", name, warps, (double)c / (iters * 16));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8}) {
    run<0>("DFMA", w); run<4>("DMUL", w); run<1>("SHFL.64 (2 x SHFL)", w); run<2>("MUFU.RSQ64H + DADD", w);
    run<3>("DMMA m8n8k4 (C chain)", w); run<5>("sqrt + DADD", w); run<6>("div + DADD", w);
  }
  return 0;
}
