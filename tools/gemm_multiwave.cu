// gemm_multiwave.cu — probe build: event-timed back-to-back launches of the residual GEMM at a multi-wave M
// (default 6000 = the C3 / C5 benches) for the slab path, the every-tile tail path (ALLTAIL) and a plain store.
// Dev tool, not shipped.   usage: gemm_multiwave [M] [bn]
#include <stdlib.h>
#include <string.h>

#include "../ace-step-1.5-for-windows_b200/csrc/epilogues.cuh"
#include "../ace-step-1.5-for-windows_b200/csrc/gemm.cuh"
using namespace ace;

template <class Epi>
static float time_it(const GemmPlan& p, const Epi& epi, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch_gemm(p, epi, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch_gemm(p, epi, 0);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel: %s / %s\n", cudaGetErrorString(e), get_error());
    exit(1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / reps;
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 6000, S = M / 2;
  const int bn = argc > 2 ? atoi(argv[2]) : -192;
  const int shapes[2][2] = {{2048, 2048}, {2048, 6144}};
  for (auto& sh : shapes) {
    const int N = sh[0], K = sh[1];
    bf16 *A, *B, *H, *G, *Gn, *Cv;
    float* Ssp;
    int* Slot;
    cudaMalloc(&A, (size_t)M * K * 2);
    cudaMalloc(&B, (size_t)N * K * 2);
    cudaMalloc(&H, (size_t)M * N * 2);
    cudaMalloc(&G, (size_t)2 * N * 2);
    cudaMalloc(&Gn, (size_t)M * N * 2);
    cudaMalloc(&Cv, (size_t)2 * N * 2);
    cudaMalloc(&Ssp, (size_t)M * 32 * 4);
    cudaMalloc(&Slot, 64);
    cudaMemset(A, 0, (size_t)M * K * 2);
    cudaMemset(B, 0, (size_t)N * K * 2);
    cudaMemset(H, 0, (size_t)M * N * 2);
    cudaMemset(G, 0, (size_t)2 * N * 2);
    cudaMemset(Cv, 0, (size_t)2 * N * 2);
    cudaMemset(Slot, 0, 64);
    GemmPlan p;
    if (make_gemm_plan(&p, A, M, K, K, B, N, K, M, 1, nullptr, bn) != ACE_OK) {
      printf("plan: %s\n", get_error());
      return 1;
    }
    NormOut no{Gn, (long)N, Cv, (long)N, Ssp, N / 64, Slot, S};
    EpiGatedResid slab{H, (long)N, G, (long)N, S, Slot, no};
    EpiGatedResid tail = slab;
    tail.use_tma = 1;
    encode_tmap_2d(&tail.tm_h, H, (uint64_t)N, (uint64_t)M, (uint64_t)N * 2, 128u);
    encode_tmap_2d(&tail.tm_g, Gn, (uint64_t)N, (uint64_t)M, (uint64_t)N * 2, 128u);
    EpiBias plain{H, (long)N, nullptr};
    const double fl = 2.0 * M * N * K;
    const float t_plain = time_it(p, plain, 20), t_slab = time_it(p, slab, 20), t_tail = time_it(p, tail, 20);
    printf("M=%d N=%d K=%d bn=%d: plain store %.1f us (%.0f TF/s) | slab %.1f us (%.0f) | every-tile tail %.1f us (%.0f)\n",
           M, N, K, p.bn, t_plain, fl / t_plain * 1e-6, t_slab, fl / t_slab * 1e-6, t_tail, fl / t_tail * 1e-6);
  }
  return 0;
}
