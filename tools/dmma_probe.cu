// DMMA issue-rate probe: TFLOP/s as a function of resident warps per SM and independent
// accumulator chains per warp.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, int iters) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
template <int CH>
void run(int warps_per_sm, double* d) {
  int threads = warps_per_sm * 32; int blocks = 148; int per = 1;
  if (threads > 1024) { threads /= 2; per = 2; blocks = 296; }
  int iters = 200000 / CH * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<blocks, threads>>>(d, 100);
  cudaEventRecord(e0); k<CH><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = (double)blocks * (threads / 32) * (double)iters * CH * 512.0;
  printf("warps/SM=%2d chains=%2d : %.2f TFLOP/s\n", warps_per_sm * 1, CH, flops / (ms * 1e-3) / 1e12);
  (void)per;
}
int main() {
  double* d; cudaMalloc(&d, 64);
  for (int w : {4, 8, 12, 16, 32}) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<16>(w, d); run<32>(w, d); }
  return 0;
}
