// tmem_probe.cu — TMEM read throughput of tcgen05.ld.32x32b.x32 per SM (4 and 8 warps), to size the
// softmax / epilogue phases: a 128 x 64 fp32 S tile is 32 KB.  Dev tool, not shipped.
#include <cstdio>
#include "../ace-step-1.5-for-windows_b200/csrc/common.cuh"
using namespace ace;

__global__ void k(float* out, int iters, int ilp) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[4][32];
    for (int u = 0; u < 4; ++u)
      if (u < ilp) tmem_ld_32x32_nowait(base + (uint32_t)(((i * 4 + u) * 32) & 511), r[u]);
    tmem_ld_wait();
    for (int u = 0; u < 4; ++u)
      if (u < ilp) acc += __uint_as_float(r[u][0]) + __uint_as_float(r[u][17]) + __uint_as_float(r[u][31]);
  }
  const long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("warps=%d loads in flight=%d: %.1f B/clk/SM (%.0f clk per 32 KB S tile)\n", blockDim.x / 32, ilp,
           (double)iters * ilp * (blockDim.x / 32) * 4096.0 / (double)(t1 - t0),
           32768.0 / ((double)iters * ilp * (blockDim.x / 32) * 4096.0 / (double)(t1 - t0)));
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 512 * 4);
  for (int threads : {128, 256})
    for (int ilp : {1, 2, 4}) {
      k<<<148, threads>>>(out, 2000, ilp);
      cudaDeviceSynchronize();
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
