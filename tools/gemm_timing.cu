// gemm_timing.cu — probe build (-DACE_GEMM_TIMING): per-phase globaltimer stamps of the CTA-pair
// GEMM for the four DiT problem shapes at the C2 bench size (M = 1500).  Dev tool, not shipped.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../ace-step-1.5-for-windows_b200/csrc/epilogues.cuh"
#include "../ace-step-1.5-for-windows_b200/csrc/gemm.cuh"
using namespace ace;

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 1500, S = M / 2;  // M = 6000: multi-wave (the every-tile tail variant)
  const int shapes[4][2] = {{2048, 2048}, {4096, 2048}, {12288, 2048}, {2048, 6144}};
  char* flush;
  cudaMalloc(&flush, 256u << 20);
  for (auto& sh : shapes) {
    const int N = sh[0], K = sh[1];
    bf16 *A, *B, *H, *G;
    cudaMalloc(&A, (size_t)M * K * 2);
    cudaMalloc(&B, (size_t)N * K * 2);
    cudaMalloc(&H, (size_t)M * N * 2);
    cudaMalloc(&G, (size_t)2 * N * 2);
    cudaMemset(A, 0, (size_t)M * K * 2);
    cudaMemset(B, 0, (size_t)N * K * 2);
    cudaMemset(H, 0, (size_t)M * N * 2);
    cudaMemset(G, 0, (size_t)2 * N * 2);
    GemmPlan p;
    if (make_gemm_plan(&p, A, M, K, K, B, N, K, M, 1, nullptr, 256) != ACE_OK) {
      printf("plan: %s\n", get_error());
      return 1;
    }
    // deferred-norm side buffers (epilogues.cuh NormOut / NormIn)
    bf16 *Gn, *Cv;
    float *Ssp, *Bs;
    int* Slot;
    cudaMalloc(&Gn, (size_t)M * N * 2);
    cudaMalloc(&Cv, (size_t)2 * N * 2);
    cudaMalloc(&Ssp, (size_t)M * 32 * 4);
    cudaMalloc(&Bs, (size_t)2 * N * 4);
    cudaMalloc(&Slot, 64);
    cudaMemset(Cv, 0, (size_t)2 * N * 2);
    cudaMemset(Ssp, 0, (size_t)M * 32 * 4);
    cudaMemset(Bs, 0, (size_t)2 * N * 4);
    cudaMemset(Slot, 0, 64);
    EpiGatedResid epi{H, (long)N, G, (long)N, S};
    EpiGatedResid epi_nogate{H, (long)N, nullptr, 0, S};
    EpiBias epi_bias{H, (long)N, nullptr};
    NormOut no{Gn, (long)N, Cv, (long)N, Ssp, N / 64, Slot, S};
    NormIn ni{Ssp, 32, 1.0f / 2048.0f, 1e-6f, Bs, (long)N, Slot, S};
    EpiGatedResid epi_no{H, (long)N, G, (long)N, S, Slot, no};
    EpiGatedResid epi_tail = epi_no;  // rep 10: the TMA tail path (residual by TMA, in-place update, TMA stores)
    epi_tail.use_tma = 1;
    encode_tmap_2d(&epi_tail.tm_h, H, (uint64_t)N, (uint64_t)M, (uint64_t)N * 2, 128u);
    encode_tmap_2d(&epi_tail.tm_g, Gn, (uint64_t)N, (uint64_t)M, (uint64_t)N * 2, 128u);
    EpiGatedResid epi_tail_plain = epi;  // rep 11: tail path without the NormOut side outputs
    epi_tail_plain.use_tma = 1;
    epi_tail_plain.tm_h = epi_tail.tm_h;
    EpiQKV epi_qkv{H, (long)N, N / 2, N / 4, G, G, nullptr, nullptr, S, 1e-6f};
    EpiQKV epi_qkv_ni{H, (long)N, N / 2, N / 4, G, G, nullptr, nullptr, S, 1e-6f, ni};
    EpiSwiGLU epi_sw{H, (long)N / 2};
    EpiSwiGLU epi_sw_ni{H, (long)N / 2, ni};
    for (int rep = 0; rep < 12; ++rep) {
      if (rep < 2) cudaMemset(flush, rep, 256u << 20);  // reps 0,1: cold L2; rep 2: warm
      cudaDeviceSynchronize();
      // rep 2: gated residual (warm); rep 3: residual without gate; rep 4: plain store (EpiBias); rep 5: gated
      // residual + NormOut; 6 / 7: EpiQKV without / with NormIn; 8 / 9: EpiSwiGLU without / with NormIn
      if (rep <= 2) launch_gemm(p, epi, 0);
      else if (rep == 3) launch_gemm(p, epi_nogate, 0);
      else if (rep == 4) launch_gemm(p, epi_bias, 0);
      else if (rep == 5) launch_gemm(p, epi_no, 0);
      else if (rep == 6) launch_gemm(p, epi_qkv, 0);
      else if (rep == 7) launch_gemm(p, epi_qkv_ni, 0);
      else if (rep == 8) launch_gemm(p, epi_sw, 0);
      else if (rep == 9) launch_gemm(p, epi_sw_ni, 0);
      else if (rep == 10) launch_gemm(p, epi_tail, 0);
      else launch_gemm(p, epi_tail_plain, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("kernel: %s\n", cudaGetErrorString(e));
        return 1;
      }
      unsigned long long st[16];
      cudaMemcpyFromSymbol(st, g_gemm_stamps, sizeof(st));
      if (rep >= 10) {
        long long tc[32];
        cudaMemcpyFromSymbol(tc, g_tail_clk, sizeof(tc));
        printf("   tail clk (cycles): box0 wait %lld | box0 math %lld | box0 fence+bar+store %lld | box1 wait %lld | box1 math %lld | "
               "box1 fence+bar+store %lld | arrive %lld | store wait %lld\n", tc[1] - tc[0], tc[2] - tc[1], tc[3] - tc[2],
               tc[5] - tc[4], tc[6] - tc[5], tc[7] - tc[6], tc[8] - tc[7], tc[9] - tc[8]);
        printf("   inside the last tail_box: issue tmem ld + gate/c ldg %lld | tmem wait %lld | chunk loop %lld\n",
               tc[17] - tc[16], tc[18] - tc[17], tc[19] - tc[18]);
      }
      printf(
          "N=%5d K=%5d rep%2d: setup %6.2f | first-issue %6.2f | first-data %6.2f | mainloop %6.2f | "
          "mma->epi %6.2f | epilogue %6.2f | teardown %6.2f | total %6.2f us\n",
          N, K, rep, (st[1] - st[0]) * 1e-3, (st[2] - st[1]) * 1e-3, (st[3] - st[2]) * 1e-3,
          (st[4] - st[3]) * 1e-3, (st[5] - st[4]) * 1e-3, (st[6] - st[5]) * 1e-3, (st[7] - st[6]) * 1e-3,
          (st[7] - st[0]) * 1e-3);
    }
    cudaFree(A);
    cudaFree(B);
    cudaFree(H);
    cudaFree(G);
  }
  return 0;
}
