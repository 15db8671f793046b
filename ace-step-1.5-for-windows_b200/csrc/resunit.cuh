// resunit.cuh — the Oobleck residual unit as ONE kernel (128-channel stages of the codec).
//
//     h  = conv7_dil(xs) + b1          xs = snake1(x), produced by the previous op's epilogue
//     hs = snake2(h)                   never leaves the SM: TMEM -> registers -> shared memory
//     x' = x + conv1(hs) + b2          second tcgen05 GEMM, A operand = hs straight from smem
//     xs'= snake_next(x')              A operand of whatever follows
//
// (diffusers AutoencoderOobleck residual unit, structure per acestep/models/mlx/vae_model.py:62-87.)
// Run as two launches of the tap-shifted GEMM this costs 4.4 GB of HBM traffic per unit at the last
// decoder stage (L = 2.88 M frames x 128 channels): hs is written and read back (1.5 GB) and the
// 1x1 convolution's epilogue is latency-bound on its residual reads (2.9 TB/s, profiles/r1_v3_*).
// Fused: 2.9 GB (read xs and x once, write x' and xs' once) and every global read goes through TMA.
//
// RU_CLUSTER > 1 makes clusters of CTAs share the conv7 weights (every 16 KB K block of W1 fetched
// once per cluster, each CTA loading a slice and multicasting it; the CTAs then step through their K
// blocks in lockstep, and a CTA whose tile index runs past the end computes on zero-filled rows).
// Measured on B200 with 4-CTA clusters: bit-identical results but 18.7 ms instead of 12.6 ms for the
// 60 s decode — a 4-deep ring cannot cover the cluster-wide release/refill round trip — so the
// shipped configuration keeps private copies (RU_CLUSTER = 1).
//
// One CTA = one 128-frame tile at a time (persistent over tiles), warp-specialised:
//   warp 0      TMA producer: per tile 14 K blocks of conv7 (7 taps x 2 halves of 64 channels; tap t
//               reads rows m0 + (t-3)*dil, out-of-range rows are zero-filled = the conv's padding),
//               plus the tile's residual rows x[m0:m0+128, :] (consumed two tiles later)
//   warp 1      MMA issuer: conv7 -> D1[it & 1] (TMEM), conv1 of the PREVIOUS tile -> D2[(it-1) & 1]
//               slotted in at K block 11 so the tensor pipe never waits for the Snake epilogue
//   warp 2      TMEM allocator (512 columns: D1 x2, D2 x2)
//   warps 4-11  epilogue 1: D1 + b1 -> bf16 -> snake2 -> bf16 -> smem in the K-major SWIZZLE_128B
//               layout the second GEMM reads (warp w: TMEM lanes 32*(w%4).., channels 64*((w-4)/4)..)
//   warps 12-19 epilogue 2: D2 + b2 -> bf16 -> + x (smem) -> x' and snake_next(x') to HBM
//   (eight warps each: with four, two resident warps per scheduler could not hide the TMEM / MUFU /
//   store latencies and the Snake math, not the operand stream, paced the kernel)
// Arithmetic and rounding points are identical to the two-launch path (EpiConv), so results match it
// bit for bit (tests/test_gpu_kernels.py::test_fused_res_unit_matches_two_launch_path).
#pragma once

#include "common.cuh"
#include "epilogues.cuh"
#include "gemm.cuh"

namespace ace {

constexpr int RU_C = 128;       // channels (in = out)
constexpr int RU_CLUSTER = 1;   // CTAs sharing each W1 K block through TMA multicast (1 = private copies, see below)
constexpr int RU_STAGES = 4;    // conv7 operand ring
constexpr int RU_THREADS = 640;  // 4 control warps + 8 (epilogue 1) + 8 (epilogue 2)
constexpr int RU_KB = 14;       // K blocks of 64 per tile: 7 taps x 2
constexpr int RU_MMA2_AT = 11;  // K block of tile it+1 after which conv1 of tile it is issued
constexpr int RU_LOADX_AT = 8;  // K block of tile it+1 after which the residual rows of tile it are fetched

struct RuSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;  // [128 rows x 64 ch] bf16, SWIZZLE_128B
  static constexpr int STAGE_BYTES = 2 * A_BYTES;
  static constexpr int TILE_BYTES = 2 * A_BYTES;  // a [128 x 128] bf16 tile = two 64-channel halves
  static constexpr int OFF_W2 = RU_STAGES * STAGE_BYTES;
  static constexpr int OFF_HS = OFF_W2 + TILE_BYTES;
  static constexpr int OFF_X = OFF_HS + TILE_BYTES;
  static constexpr int OFF_BAR = OFF_X + TILE_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;  // +1024: manual alignment
};

struct RuParams {
  int L;    // frames (rows)
  int dil;  // dilation of the 7-tap convolution
  const float *bias1, *a2, *ib2;   // conv7 bias, snake2 exp(alpha), 1/(exp(beta)+1e-9)
  const float *bias2, *an, *ibn;   // conv1 bias, next op's Snake
  bf16 *ox, *oxs;                  // x' and snake_next(x'), [L, 128]
};

#ifdef __CUDACC__

enum RuBar {
  RU_FULL = 0, RU_EMPTY = RU_STAGES, RU_W2 = 2 * RU_STAGES, RU_D1F, RU_D1F1, RU_D1E, RU_D1E1, RU_HSF, RU_HSE,
  RU_D2F, RU_D2F1, RU_D2E, RU_D2E1, RU_XF, RU_XE, RU_NBAR
};

__global__ void __cluster_dims__(RU_CLUSTER, 1, 1) __launch_bounds__(RU_THREADS, 1)
res_unit_kernel(const __grid_constant__ CUtensorMap tm_xs, const __grid_constant__ CUtensorMap tm_w1,
                const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_x,
                const RuParams p) {
  static_assert(128 % RU_CLUSTER == 0 && (128 / RU_CLUSTER) % 8 == 0, "W1 slices must be whole swizzle atoms");
  constexpr int W1_ROWS = 128 / RU_CLUSTER;          // rows of each W1 K block this CTA fetches
  constexpr uint16_t ALL = (1u << RU_CLUSTER) - 1;   // multicast mask: every CTA of the cluster
  using S = RuSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sW2 = smem + S::OFF_W2;
  uint8_t* sHS = smem + S::OFF_HS;
  uint8_t* sX = smem + S::OFF_X;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + RU_NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_xs);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < RU_STAGES; ++i) {
      mbar_init(&bar[RU_FULL + i], 1);
      mbar_init(&bar[RU_EMPTY + i], RU_CLUSTER);  // the slot is rewritten by every CTA's multicast
    }
    mbar_init(&bar[RU_W2], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar[RU_D1F + i], 1);
      mbar_init(&bar[RU_D1E + i], 8);
      mbar_init(&bar[RU_D2F + i], 1);
      mbar_init(&bar[RU_D2E + i], 8);
    }
    mbar_init(&bar[RU_HSF], 8);
    mbar_init(&bar[RU_HSE], 1);
    mbar_init(&bar[RU_XF], 1);
    mbar_init(&bar[RU_XE], 8);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  cluster_sync_all();  // barriers of every CTA are initialised before any remote arrive / multicast
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = cluster_ctarank();

  // every CTA runs the same number of tile iterations (lockstep on the shared W1 stream); tiles past
  // the end read zero-filled rows and store nothing
  const int num_tiles = (p.L + 127) / 128;
  const int my_tiles = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    const bool elected = elect_one();
    if (elected) {  // conv1 weights: constant, so they may be fetched before the dependency wait
      mbar_arrive_expect_tx(&bar[RU_W2], S::TILE_BYTES);
      tma_load_2d(sW2, &tm_w2, &bar[RU_W2], 0, 0);
      tma_load_2d(sW2 + S::A_BYTES, &tm_w2, &bar[RU_W2], 64, 0);
    }
    pdl_wait();
    auto load_x = [&](int j) {  // residual rows of this CTA's j-th tile (single buffer)
      mbar_wait(&bar[RU_XE], (uint32_t)((j & 1) ^ 1));
      if (elected) {
        const int m0 = ((int)blockIdx.x + j * (int)gridDim.x) * 128;
        mbar_arrive_expect_tx(&bar[RU_XF], S::TILE_BYTES);
        tma_load_2d(sX, &tm_x, &bar[RU_XF], 0, m0);
        tma_load_2d(sX + S::A_BYTES, &tm_x, &bar[RU_XF], 64, m0);
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128;
      for (int kb = 0; kb < RU_KB; ++kb) {
        const int tap = kb >> 1, kk = kb & 1;
        mbar_wait(&bar[RU_EMPTY + stage], phase ^ 1);
        if (elected) {
          mbar_arrive_expect_tx(&bar[RU_FULL + stage], S::STAGE_BYTES);
          tma_load_2d(smem + stage * S::STAGE_BYTES, &tm_xs, &bar[RU_FULL + stage], kk * 64,
                      m0 + (tap - 3) * p.dil);
          if (RU_CLUSTER == 1) {
            tma_load_2d(smem + stage * S::STAGE_BYTES + S::A_BYTES, &tm_w1, &bar[RU_FULL + stage],
                        tap * RU_C + kk * 64, 0);
          } else {
            tma_load_2d_mcast(smem + stage * S::STAGE_BYTES + S::A_BYTES + crank * (W1_ROWS * 128), &tm_w1,
                              &bar[RU_FULL + stage], tap * RU_C + kk * 64, (int)crank * W1_ROWS, ALL);
          }
        }
        if (++stage == RU_STAGES) {
          stage = 0;
          phase ^= 1;
        }
        // the previous tile's residual rows are needed only after its conv1, which the MMA warp issues
        // at K block RU_MMA2_AT of this tile; by now epilogue 2 of the tile before that has long
        // released the buffer, so this wait never stalls the loads that feed the tensor core
        if (kb == RU_LOADX_AT && it > 0) load_x(it - 1);
      }
    }
    if (my_tiles > 0) load_x(my_tiles - 1);
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_umma_idesc_bf16(128, RU_C);
    const bool elected = elect_one();
    const uint32_t ring_lo = (smem_u32(smem) >> 4) & 0x3FFFu;
    const uint32_t hs_lo = (smem_u32(sHS) >> 4) & 0x3FFFu, w2_lo = (smem_u32(sW2) >> 4) & 0x3FFFu;
    auto mma2 = [&](int j) {  // x'_j accumulators: D2[j & 1] = hs_j . W2^T   (K = 128 = 2 halves x 4 slices)
      const int b = j & 1;
      mbar_wait(&bar[RU_HSF], (uint32_t)(j & 1));
      mbar_wait(&bar[RU_D2E + b], (uint32_t)(((j >> 1) & 1) ^ 1));
      tcgen05_fence_after();
      if (elected) {
        const uint32_t d = tmem_base + 256u + (uint32_t)(b * RU_C);
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_lo(d, hs_lo + (uint32_t)(half * (S::A_BYTES >> 4) + 2 * k),
                            w2_lo + (uint32_t)(half * (S::A_BYTES >> 4) + 2 * k), idesc, (half | k) != 0 ? 1u : 0u);
        umma_commit(&bar[RU_HSE]);
        umma_commit(&bar[RU_D2F + b]);
      }
    };
    mbar_wait(&bar[RU_W2], 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      mbar_wait(&bar[RU_D1E + b], (uint32_t)(((it >> 1) & 1) ^ 1));
      tcgen05_fence_after();
      const uint32_t d1 = tmem_base + (uint32_t)(b * RU_C);
      for (int kb = 0; kb < RU_KB; ++kb) {
        mbar_wait(&bar[RU_FULL + stage], phase);
        tcgen05_fence_after();
        if (elected) {
          const uint32_t a_lo = ring_lo + (uint32_t)stage * (S::STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + (S::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo(d1, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          if (RU_CLUSTER == 1) {
            umma_commit(&bar[RU_EMPTY + stage]);
          } else {
            umma_commit_mcast(&bar[RU_EMPTY + stage], ALL);  // frees the slot in every CTA that multicasts into it
          }
          if (kb == RU_KB - 1) umma_commit(&bar[RU_D1F + b]);
        }
        if (++stage == RU_STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (kb == RU_MMA2_AT && it > 0) mma2(it - 1);
      }
    }
    if (my_tiles > 0) mma2(my_tiles - 1);
  } else if (warp >= 4 && warp < 12) {
    // ---------------- epilogue 1: D1 -> snake2 -> hs tile in smem (A operand of conv1) ----------------
    const int quarter = (warp - 4) & 3, r = quarter * 32 + lane;
    const int c_lo = ((warp - 4) >> 2) * 64;  // this warp's 64 channels
    uint8_t* rowp = sHS + r * 128;
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      mbar_wait(&bar[RU_D1F + b], (uint32_t)((it >> 1) & 1));
      mbar_wait(&bar[RU_HSE], (uint32_t)((it & 1) ^ 1));  // conv1 of the previous tile has read hs
      tcgen05_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * RU_C);
#pragma unroll 1
      for (int c = c_lo; c < c_lo + 64; c += 32) {
        float v[32];
        tmem_ld_32x32(taddr + (uint32_t)c, v);
        // bf16 conversions run on the XU pipe (16 per clock per SM, shared with the Snake cosine) and were
        // what paced this kernel: every rounding is therefore done two elements at a time (F2FP.PACK_AB)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias1 + c) + i);
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.a2 + c) + i);
          const float4 i4 = __ldg(reinterpret_cast<const float4*>(p.ib2 + c) + i);
          float h0, h1, h2, h3;
          unpack_bf16x2(pack_bf16x2(v[4 * i + 0] + b4.x, v[4 * i + 1] + b4.y), h0, h1);
          unpack_bf16x2(pack_bf16x2(v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w), h2, h3);
          v[4 * i + 0] = snake_f(h0, a4.x, i4.x);
          v[4 * i + 1] = snake_f(h1, a4.y, i4.y);
          v[4 * i + 2] = snake_f(h2, a4.z, i4.z);
          v[4 * i + 3] = snake_f(h3, a4.w, i4.w);
        }
        uint8_t* half = rowp + (c >> 6) * S::A_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (((c & 32) >> 3) + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(half + chunk * 16) =
              make_uint4(pack_bf16x2(v[8 * q + 0], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                         pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar[RU_D1E + b]);
        mbar_arrive(&bar[RU_HSF]);
      }
    }
  } else if (warp >= 12) {
    // ---------------- epilogue 2: D2 + b2 + x -> x', snake_next(x') ----------------
    const int quarter = (warp - 12) & 3, r = quarter * 32 + lane;
    const int c_lo = ((warp - 12) >> 2) * 64;  // this warp's 64 channels
    const uint8_t* xrow = sX + r * 128;
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const long row = ((long)blockIdx.x + (long)it * gridDim.x) * 128 + r;
      mbar_wait(&bar[RU_D2F + b], (uint32_t)((it >> 1) & 1));
      mbar_wait(&bar[RU_XF], (uint32_t)(it & 1));
      tcgen05_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + (uint32_t)(b * RU_C);
      const bool ok = row < p.L;
#pragma unroll 1
      for (int c = c_lo; c < c_lo + 64; c += 32) {
        float v[32];
        tmem_ld_32x32(taddr + (uint32_t)c, v);
        const uint8_t* half = xrow + (c >> 6) * S::A_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (((c & 32) >> 3) + q) ^ (r & 7);
          const uint4 xq = *reinterpret_cast<const uint4*>(half + chunk * 16);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias2 + c) + 2 * q);
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias2 + c) + 2 * q + 1);
          // x' = bf16(bf16(acc + b2) + x): one pair-wise rounding, then a packed bf16 add (exactly rounded)
          uint4 xn;
          xn.x = badd2(pack_bf16x2(v[8 * q + 0] + b0.x, v[8 * q + 1] + b0.y), xq.x);
          xn.y = badd2(pack_bf16x2(v[8 * q + 2] + b0.z, v[8 * q + 3] + b0.w), xq.y);
          xn.z = badd2(pack_bf16x2(v[8 * q + 4] + b1.x, v[8 * q + 5] + b1.y), xq.z);
          xn.w = badd2(pack_bf16x2(v[8 * q + 6] + b1.z, v[8 * q + 7] + b1.w), xq.w);
          if (ok) reinterpret_cast<uint4*>(p.ox + row * RU_C + c)[q] = xn;
          unpack_bf16x2(xn.x, v[8 * q + 0], v[8 * q + 1]); unpack_bf16x2(xn.y, v[8 * q + 2], v[8 * q + 3]);
          unpack_bf16x2(xn.z, v[8 * q + 4], v[8 * q + 5]); unpack_bf16x2(xn.w, v[8 * q + 6], v[8 * q + 7]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.an + c) + i);
          const float4 i4 = __ldg(reinterpret_cast<const float4*>(p.ibn + c) + i);
          v[4 * i + 0] = snake_f(v[4 * i + 0], a4.x, i4.x);
          v[4 * i + 1] = snake_f(v[4 * i + 1], a4.y, i4.y);
          v[4 * i + 2] = snake_f(v[4 * i + 2], a4.z, i4.z);
          v[4 * i + 3] = snake_f(v[4 * i + 3], a4.w, i4.w);
        }
        if (ok) store_bf16x32(p.oxs + row * RU_C + c, v);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar[RU_D2E + b]);
        mbar_arrive(&bar[RU_XE]);
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();  // no CTA may exit while a sibling can still multicast into it
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// xs, x: [L, 128] bf16 (xs = snake1(x)); w1: [128, 7*128] tap-major; w2: [128, 128]; outputs [L, 128].
inline int launch_res_unit_fused(const bf16* xs, const bf16* x, const bf16* w1, const bf16* w2, const RuParams& p,
                                 cudaStream_t stream) {
  if (p.L <= 0) return ACE_OK;
  CUtensorMap tm_xs, tm_w1, tm_w2, tm_x;
  ACE_PROPAGATE(encode_tmap_2d(&tm_xs, xs, RU_C, (uint64_t)p.L, RU_C * sizeof(bf16), 128));
  ACE_PROPAGATE(encode_tmap_2d(&tm_w1, w1, 7 * RU_C, RU_C, 7 * RU_C * sizeof(bf16), 128 / RU_CLUSTER));
  ACE_PROPAGATE(encode_tmap_2d(&tm_w2, w2, RU_C, RU_C, RU_C * sizeof(bf16), 128));
  ACE_PROPAGATE(encode_tmap_2d(&tm_x, x, RU_C, (uint64_t)p.L, RU_C * sizeof(bf16), 128));
  static bool attr_set = false;
  if (!attr_set) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(res_unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RuSmem::TOTAL));
    attr_set = true;
  }
  const int tiles = (p.L + 127) / 128;
  const int max_grid = num_sms() / RU_CLUSTER * RU_CLUSTER;
  int grid = (tiles + RU_CLUSTER - 1) / RU_CLUSTER * RU_CLUSTER;
  if (grid > max_grid) grid = max_grid;
  prof_tag_gemm(p.L, RU_C, 8 * RU_C);
  prof_begin(PROF_GEMM, 2.0 * p.L * RU_C * 8.0 * RU_C, 4.0 * p.L * RU_C * 2.0, stream);
  ACE_CUDA_CHECK(launch_kernel(res_unit_kernel, dim3(grid), dim3(RU_THREADS), (size_t)RuSmem::TOTAL, stream, tm_xs,
                               tm_w1, tm_w2, tm_x, p));
  prof_end(stream);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

#endif  // __CUDACC__

}  // namespace ace
