// ru_timing.cu — probe build (-DACE_RU_TIMING): what the MMA warp of the fused residual-unit kernel waits for.
// Dev tool, not shipped.   usage: ru_timing [frames] [dil]
#include <cstdio>
#include <cstdlib>

#include "../ace-step-1.5-for-windows_b200/csrc/resunit.cuh"
using namespace ace;

int main(int argc, char** argv) {
  const long L = argc > 1 ? atol(argv[1]) : 2880000;
  const int dil = argc > 2 ? atoi(argv[2]) : 9;
  bf16 *xs, *x, *w1, *w2, *ox, *oxs;
  float* vec;
  cudaMalloc(&xs, L * 128 * 2); cudaMalloc(&x, L * 128 * 2); cudaMalloc(&ox, L * 128 * 2); cudaMalloc(&oxs, L * 128 * 2);
  cudaMalloc(&w1, 128 * 7 * 128 * 2); cudaMalloc(&w2, 128 * 128 * 2); cudaMalloc(&vec, 6 * 128 * 4);
  cudaMemset(xs, 0, L * 128 * 2); cudaMemset(x, 0, L * 128 * 2);
  cudaMemset(w1, 0, 128 * 7 * 128 * 2); cudaMemset(w2, 0, 128 * 128 * 2); cudaMemset(vec, 0, 6 * 128 * 4);
  RuParams p{(int)L, dil, vec, vec + 128, vec + 256, vec + 384, vec + 512, vec + 640, ox, oxs};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (launch_res_unit_fused(xs, x, w1, w2, p, 0) != ACE_OK) { printf("launch: %s\n", get_error()); return 1; }
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long w[8];
    cudaMemcpyFromSymbol(w, g_ru_wait, sizeof(w));
    long long e2[8];
    cudaMemcpyFromSymbol(e2, g_ru_e2, sizeof(e2));
    if (rep == 2)
      printf("   epilogue 2 (warp 12), cycles per tile: total %.0f | wait D2 %.0f | wait x tile %.0f | TMEM + math %.0f | x' out %.0f | xs' out %.0f\n",
             e2[5] / (double)w[6], e2[0] / (double)w[6], e2[1] / (double)w[6], e2[2] / (double)w[6], e2[3] / (double)w[6], e2[4] / (double)w[6]);
    long long e1[8];
    cudaMemcpyFromSymbol(e1, g_ru_e1, sizeof(e1));
    if (rep == 2)
      printf("   epilogue 1 (warp 4), cycles per tile: total %.0f | wait D1 %.0f | wait hs free %.0f | TMEM + snake + hs store %.0f\n",
             e1[3] / (double)w[6], e1[0] / (double)w[6], e1[1] / (double)w[6], e1[2] / (double)w[6]);
    long long pw[8];
    cudaMemcpyFromSymbol(pw, g_ru_pwait, sizeof(pw));
    const double t = (double)w[6];
    if (rep == 2)
      printf("   producer warp, cycles per tile: total %.0f | waits: ring slot free %.0f | xs buffer free %.0f | x buffer free %.0f\n",
             pw[3] / t, pw[0] / t, pw[1] / t, pw[2] / t);
    if (rep == 2)
      printf("L=%ld dil=%d: %.1f us | MMA warp of CTA 0, cycles per tile over %lld tiles: total %.0f | waits: weights %.0f | xs box %.0f | "
             "D1 free (epilogue 1) %.0f | hs ready (epilogue 1) %.0f | D2 free (epilogue 2) %.0f\n", L, dil, ms * 1e3, w[6],
             w[5] / t, w[0] / t, w[1] / t, w[2] / t, w[3] / t, w[4] / t);
  }
  return 0;
}
