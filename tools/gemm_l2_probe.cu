// gemm_l2_probe.cu — probe build (-DACE_GEMM_TIMING): is the CTA-pair GEMM main loop bound by the
// chip-wide L2 -> SM bandwidth or by a per-SM limit?  Runs one large problem on 74 / 56 / 37 / 18 / 8
// clusters and prints SM cycles per 64-deep K block on cluster 0 (tensor-core floor: 512 cycles for a
// 256 x 256 pair tile, 384 for 256 x 192).  If cycles per K block fall as fewer SMs run, the shared
// L2 is the limiter; if they stay put, the limit is per SM.  Dev tool, not shipped.
#include <stdlib.h>
#include <string.h>

#include "../ace-step-1.5-for-windows_b200/csrc/epilogues.cuh"
#include "../ace-step-1.5-for-windows_b200/csrc/gemm.cuh"
using namespace ace;

int main() {
  const int M = 6144, N = 12288, K = 2048;
  bf16 *A, *B, *H, *G;
  cudaMalloc(&A, (size_t)M * K * 2);
  cudaMalloc(&B, (size_t)N * K * 2);
  cudaMalloc(&H, (size_t)M * N * 2);
  cudaMalloc(&G, (size_t)2 * N * 2);
  cudaMemset(A, 0, (size_t)M * K * 2);
  cudaMemset(B, 0, (size_t)N * K * 2);
  cudaMemset(H, 0, (size_t)M * N * 2);
  cudaMemset(G, 0, (size_t)2 * N * 2);
  for (int bn : {256, 192}) {
    GemmPlan p;
    if (make_gemm_plan(&p, A, M, K, K, B, N, K, M, 1, nullptr, bn) != ACE_OK) {
      printf("plan: %s\n", get_error());
      return 1;
    }
    EpiGatedResid epi{H, (long)N, G, (long)N, M / 2};
    for (int clusters : {74, 37, 8}) {
      char buf[16];
      snprintf(buf, sizeof buf, "%d", clusters);
      setenv("ACE_GEMM_MAX_CLUSTERS", buf, 1);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      launch_gemm(p, epi, 0);
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) launch_gemm(p, epi, 0);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("kernel: %s\n", cudaGetErrorString(e));
        return 1;
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 3;
      long long cyc[4];
      cudaMemcpyFromSymbol(cyc, g_gemm_cycles, sizeof(cyc));
      const double tf = 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12;
      printf("bn=%d clusters=%2d: %8.1f us  %7.1f TFLOP/s  %6.2f TFLOP/s/SM  | cluster 0: %lld k-blocks, %.1f cycles/k-block\n",
             bn, clusters, ms * 1e3, tf, tf / (2 * clusters), cyc[2], (double)(cyc[1] - cyc[0]) / (double)cyc[2]);
    }
  }
  return 0;
}
