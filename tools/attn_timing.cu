// attn_timing.cu — probe build (-DACE_ATTN_TIMING): where one softmax warp of the tcgen05 flash-attention kernel
// spends its cycles per KV block (C3 shape: 3000 queries x 3000 keys, 16 / 8 heads, batch 2).  Dev tool.
#include <cstdio>
#include <vector>
#include "../ace-step-1.5-for-windows_b200/csrc/attention_tc.cu"
using namespace ace;

int main() {
  const int B = 2, H = 16, HK = 8, S = 3000;
  bf16 *q, *k, *v, *o;
  cudaMalloc(&q, (size_t)B * S * H * 128 * 2);
  cudaMalloc(&k, (size_t)B * S * HK * 128 * 2);
  cudaMalloc(&v, (size_t)B * S * HK * 128 * 2);
  cudaMalloc(&o, (size_t)B * S * H * 128 * 2);
  cudaMemset(q, 0, (size_t)B * S * H * 128 * 2);
  cudaMemset(k, 0, (size_t)B * S * HK * 128 * 2);
  cudaMemset(v, 0, (size_t)B * S * HK * 128 * 2);
  for (int window : {-1, 128}) {
    AttnParams p{q, k, v, o, H * 128L, HK * 128L, HK * 128L, H * 128L, S, S, window, H / HK,
                 (1.0f / sqrtf(128.0f)) * 1.4426950408889634f, nullptr};
    AttnPlan plan;
    if (make_attn_plan(&plan, p, H, B) != ACE_OK) { printf("plan: %s\n", get_error()); return 1; }
    launch_attention_tc(plan, 0);
    cudaDeviceSynchronize();
    long long z[8] = {0};
    cudaMemcpyToSymbol(g_attn_cycles, z, sizeof(z));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch_attention_tc(plan, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c[8];
    cudaMemcpyFromSymbol(c, g_attn_cycles, sizeof(c));
    const int nblk = window < 0 ? (S + 63) / 64 : 4;  // CTA 0 (q0 = 0): band = keys [0, 256)
    const double keys = window < 0 ? S : 257;
    printf("window %4d: %.1f us (%.0f TFLOP/s) | CTA 0 warp 4, cycles per KV block over %d blocks: wait S %.0f | tmem ld %.0f | "
           "max %.0f | exp %.0f | wait PV %.0f | store P %.0f | total %.0f\n", window, ms * 1e3,
           4.0 * S * keys * 128 * H * B / (ms * 1e-3) / 1e12, nblk, (double)c[0] / nblk, (double)c[1] / nblk,
           (double)c[2] / nblk, (double)c[3] / nblk, (double)c[4] / nblk, (double)c[5] / nblk,
           (double)(c[0] + c[1] + c[2] + c[3] + c[4] + c[5]) / nblk);
  }
  // C2 shapes (750 tokens, batch 2): the fixed cost of a launch.  globaltimer stamps of CTA (0,0,0).
  {
    const int S2 = 750, E2 = 512;
    struct { const char* name; int skv, window; } cases[3] = {{"self full", S2, -1}, {"self window", S2, 128}, {"cross", E2, -1}};
    for (auto& cs : cases) {
      AttnParams p{q, k, v, o, H * 128L, HK * 128L, HK * 128L, H * 128L, S2, cs.skv, cs.window, H / HK,
                   (1.0f / sqrtf(128.0f)) * 1.4426950408889634f, nullptr};
      AttnPlan plan;
      if (make_attn_plan(&plan, p, H, B) != ACE_OK) { printf("plan: %s\n", get_error()); return 1; }
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        launch_attention_tc(plan, 0);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long st[8];
        cudaMemcpyFromSymbol(st, g_attn_stamp, sizeof(st));
        if (rep == 2)
          printf("%-12s %.1f us by events | CTA 0 (ns): setup %llu | Q,K0 load + S_0 %llu | KV loop %llu | last PV %llu | O readout + stores %llu | "
                 "teardown %llu | entry->exit %llu\n", cs.name, ms * 1e3, st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3],
                 st[5] - st[4], st[6] - st[5], st[6] - st[0]);
      }
    }
  }
  return 0;
}
