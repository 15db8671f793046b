// desc_rowstep_probe.cu — does a K-major SWIZZLE_128B UMMA descriptor accept a start address that is NOT on an
// 8-row (1024 B) boundary when its base-offset field (bits [49,52)) carries (address >> 7) & 7?  If so, the 7 taps
// of a dilated convolution can read ONE [128 + 6 dil rows] A box at row offsets instead of 7 shifted boxes.
// A: [192 rows x 64] bf16 loaded by TMA (SWIZZLE_128B) as one box; B: [128 x 64]; D_r = A[r : r+128] . B^T for
// several r, compared on the host.  Dev tool, not shipped.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ace-step-1.5-for-windows_b200/csrc/gemm.cuh"
using namespace ace;

__device__ __forceinline__ void mma_desc(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* out, int r, int use_bo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;               // 192 rows x 128 B = 24 KB
  uint8_t* sB = smem + 24576;       // 128 rows x 128 B = 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24576 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(slot, 128);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], 24576 + 16384);
    tma_load_2d(sA, &tm_a, &bar[0], 0, 0);
    tma_load_2d(sB, &tm_b, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tcgen05_fence_after();
    const uint32_t idesc = make_umma_idesc_bf16(128, 128);
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)r * 128u;
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_umma_desc_k128(a_addr + 32u * k);
      if (use_bo) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
      const uint64_t db = make_umma_desc_k128(smem_u32(sB) + 32u * k);
      mma_desc(tmem, da, db, idesc, k != 0);
    }
    umma_commit(&bar[1]);
  }
  __syncwarp();
  mbar_wait(&bar[1], 0);
  tcgen05_fence_after();
  __syncwarp();
  for (int c = 0; c < 128; c += 32) {
    float v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 128 + c + i] = v[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

static float bf(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  const int RA = 192, K = 64, N = 128;
  std::vector<uint16_t> hA(RA * K), hB(N * K);
  srand(1);
  for (auto& x : hA) x = (uint16_t)(0x3C00 + (rand() % 512));  // bf16 values around 0.008..2
  for (auto& x : hB) x = (uint16_t)(0x3C00 + (rand() % 512));
  bf16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, 128 * 128 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (encode_tmap_2d(&ta, dA, K, RA, K * 2, RA) != ACE_OK || encode_tmap_2d(&tb, dB, K, N, K * 2, N) != ACE_OK) {
    printf("tmap: %s\n", get_error());
    return 1;
  }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576 + 16384 + 1024 + 256);
  std::vector<float> hO(128 * 128);
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int r : {0, 8, 1, 3, 9, 27, 54, 63}) {
      probe<<<1, 128, 24576 + 16384 + 1024 + 256>>>(ta, tb, dO, r, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("r=%d: %s\n", r, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)bf(hA[(m + r) * K + k]) * bf(hB[n * K + k]);
          maxerr = fmax(maxerr, fabs(ref - hO[m * 128 + n]));
        }
      printf("base_offset field %s, row offset %2d: max |err| %.3g %s\n", use_bo ? "set " : "zero", r, maxerr,
             maxerr < 1e-2 ? "OK" : "WRONG");
    }
  return 0;
}
