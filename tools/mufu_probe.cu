// mufu_probe.cu — does ex2.approx.f16x2 (two exponentials per MUFU op) run at the same instruction
// rate as ex2.approx.ftz.f32 on sm_100a?  Prints exponentials per clock per SM for both.  Dev tool.
#include <cstdio>
#include <cuda_fp16.h>
__global__ void k_f32(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    a -= 1.f; b -= 1.f; c -= 1.f; d -= 1.f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("f32   : %.2f exp/clk/SM (%d threads/SM)\n", 4.0 * iters * blockDim.x / (double)(t1 - t0), blockDim.x);
}
__global__ void k_f16x2(float* out, int iters) {
  unsigned a = 0x3c003800u + threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(c));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(d));
    a ^= 0x00010001u; b ^= 0x00010001u; c ^= 0x00010001u; d ^= 0x00010001u;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(a ^ b ^ c ^ d);
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("f16x2 : %.2f exp/clk/SM (%d threads/SM)\n", 8.0 * iters * blockDim.x / (double)(t1 - t0), blockDim.x);
}
__global__ void k_bf16x2(float* out, int iters) {
  unsigned a = 0x3f803f00u + threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a));
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b));
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(c));
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(d));
    a ^= 0x00010001u; b ^= 0x00010001u; c ^= 0x00010001u; d ^= 0x00010001u;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(a ^ b ^ c ^ d);
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("bf16x2: %.2f exp/clk/SM (%d threads/SM)\n", 8.0 * iters * blockDim.x / (double)(t1 - t0), blockDim.x);
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  for (int threads : {128, 512, 1024}) {
    k_f32<<<148, threads>>>(out, 4096);
    k_f16x2<<<148, threads>>>(out, 4096);
    k_bf16x2<<<148, threads>>>(out, 4096);
    cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
