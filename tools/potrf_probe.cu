// Cycle-level breakdown of the diagonal-tile kernel (k_potrf128c), development aid:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DVC_POTRF_CLOCKS -Iinclude -Ivarycoef_b200/csrc \
//        tools/potrf_probe.cu -o tools/bin/potrf_probe && tools/bin/potrf_probe
#include "../varycoef_b200/csrc/vc_chol.cu"
#include <cstdio>
#include <vector>
#include <cmath>

int main() {
  const int n = TILE;
  std::vector<double> A(n * n), L(n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[i + n * j] = exp(-fabs((double)(i - j)) / 7.0) + (i == j ? 0.3 : 0.0);
  // host Cholesky
  L = A;
  for (int j = 0; j < n; ++j) {
    double s = L[j + n * j];
    for (int k = 0; k < j; ++k) s -= L[j + n * k] * L[j + n * k];
    L[j + n * j] = sqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = L[i + n * j];
      for (int k = 0; k < j; ++k) t -= L[i + n * k] * L[j + n * k];
      L[i + n * j] = t / L[j + n * j];
    }
  }
  double *dA, *dlinv, *dlog;
  int* dinfo;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dlinv, n * n * 8); cudaMemset(dlinv, 0, n * n * 8); cudaMalloc(&dlog, 64); cudaMalloc(&dinfo, 4);
  int big = INT_MAX;
  cudaFuncSetAttribute(k_potrf128c<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRFC_SMEM);
  cudaFuncSetAttribute(k_potrf128b<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRFB_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int variant = 0; variant < 2; ++variant) {
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(dinfo, &big, 4, cudaMemcpyHostToDevice);
      cudaEventRecord(e0);
      if (variant == 0) k_potrf128c<true><<<1, POTRF_THREADS, POTRFC_SMEM>>>(dA, n, 0, dlinv, dlog, dinfo);
      else k_potrf128b<true><<<1, POTRF_THREADS, POTRFB_SMEM>>>(dA, n, 0, dlinv, dlog, dinfo);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%s rep %d: %.2f us (%s)\n", variant == 0 ? "potrf128c" : "potrf128b", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<double> R(n * n), X(n * n);
    cudaMemcpy(R.data(), dA, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(X.data(), dlinv, n * n * 8, cudaMemcpyDeviceToHost);
    double eL = 0, eX = 0;
    for (int j = 0; j < n; ++j)
      for (int i = j; i < n; ++i) eL = fmax(eL, fabs(R[i + n * j] - L[i + n * j]));
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {   // (L X)(i, j) = delta
        double s = 0;
        for (int k = 0; k < n; ++k) s += (i >= k ? L[i + n * k] : 0.0) * X[k + n * j];
        eX = fmax(eX, fabs(s - (i == j ? 1.0 : 0.0)));
      }
    printf("  max |L - L_host| = %.2e   max |L X - I| = %.2e\n", eL, eX);
    if (variant == 0) {
      long long clk[64];
      cudaMemcpyFromSymbol(clk, g_potrf_clk, sizeof(clk));
      printf("  cycles: load %lld", clk[1] - clk[0]);
      long long p1 = 0, p2 = 0, p2a = 0, p3 = 0;
      for (int kb = 0; kb < 8; ++kb) {
        const long long prev = kb == 0 ? clk[1] : clk[5 + 4 * (kb - 1)];
        p1 += clk[2 + 4 * kb] - prev; p2a += clk[3 + 4 * kb] - clk[2 + 4 * kb];
        p2 += clk[4 + 4 * kb] - clk[2 + 4 * kb]; p3 += clk[5 + 4 * kb] - clk[4 + 4 * kb];
        printf("\n    kb %d: phase1 %lld  diag-block(warp0) %lld  phase2 %lld  phase3 %lld", kb, clk[2 + 4 * kb] - prev,
               clk[3 + 4 * kb] - clk[2 + 4 * kb], clk[4 + 4 * kb] - clk[2 + 4 * kb], clk[5 + 4 * kb] - clk[4 + 4 * kb]);
      }
      printf("\n  last diag block: load->factor start %lld, factor %lld, inverse %lld, store %lld", clk[40] - clk[30], clk[41] - clk[40], clk[42] - clk[41], clk[31] - clk[42]);
      printf("\n  totals: phase1 %lld  diag %lld  phase2 %lld  phase3 %lld  tail-xrow %lld  storeL %lld  store-inv %lld  total %lld\n",
             p1, p2a, p2, p3, clk[34] - clk[33], clk[35] - clk[34], clk[36] - clk[35], clk[36] - clk[0]);
    }
  }
  return 0;
}
