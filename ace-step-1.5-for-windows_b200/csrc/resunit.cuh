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
// Operand traffic into shared memory per 128-frame tile: the seven taps read ONE box of xs —
// rows [m0 - 3 dil, m0 + 128 + 3 dil), fetched once per 64-channel half — through UMMA descriptors whose start
// address is stepped by tap * dil rows (a K-major SWIZZLE_128B descriptor may start on any 128-byte row: the swizzle
// follows the absolute shared-memory address, base-offset field zero; tools/desc_rowstep_probe.cu).  Seven shifted
// [128 x 128] boxes were 224 KB per tile; the one box is 34-47 KB.  The weights (W1: 14 blocks of 16 KB, W2: 2) stream
// through a 4-stage ring per tile: 303-335 KB per tile in all, against 480 KB before.  What bounds the kernel is
// epilogue 2 (tools/ru_timing.cu: ~10.5k busy cycles per tile against 4.1k of tensor work), so the shared memory goes
// where it keeps THAT warp group fed: two residual tiles (the next one is fetched while epilogue 2 still copies the
// previous one out) and a single xs box (the tensor pipe idles for its round trip once per tile, for free).  (Sharing the W1 stream
// between the CTAs of a cluster by TMA multicast was measured twice — 4-CTA clusters: 18.7 vs 12.6 ms per 60 s
// decode, 2-CTA: within 1 % — multicast saves L2 reads, not the bytes each SM has to take in; it is gone.)
//
// One CTA = one 128-frame tile at a time (persistent over tiles), warp-specialised:
//   warp 0      TMA producer: the xs box of tile it (ONE buffer, requested behind the tile's first weight blocks once
//               the previous tile's conv7 has retired; out-of-range rows are zero-filled = the conv's padding), the
//               14 W1 blocks of tile it with the 2 W2 blocks of tile it-1's conv1 slotted in after block
//               RU_MMA2_AT, and the tile's residual rows x[m0:m0+128, :] into one of TWO buffers
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
constexpr int RU_STAGES = 4;    // weight ring (16 KB blocks)
constexpr int RU_THREADS = 640;  // 4 control warps + 8 (epilogue 1) + 8 (epilogue 2)
constexpr int RU_KB = 14;       // K blocks of 64 per tile: 7 taps x 2
constexpr int RU_MMA2_AT = 11;  // K block of tile it+1 after which conv1 of tile it is issued
constexpr int RU_LOADX_AT = 8;  // K block of tile it+1 after which the residual rows of tile it are fetched
constexpr int RU_LOADA_AT = 2;  // weight block of tile it after which the xs box of tile it is requested (see below)
constexpr int RU_MAX_DIL = 9;

struct RuSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;  // [128 rows x 64 ch] bf16, SWIZZLE_128B
  static constexpr int TILE_BYTES = 2 * A_BYTES;  // a [128 x 128] bf16 tile = two 64-channel halves
  static constexpr int XS_ROWS = (128 + 6 * RU_MAX_DIL + 7) / 8 * 8;  // 184: rows of the largest xs box, whole atoms
  static constexpr int XS_HALF = XS_ROWS * 128;                        // one 64-channel half of the box
  static constexpr int XS_BYTES = 2 * XS_HALF;
  static constexpr int OFF_W = XS_BYTES;                               // after the xs box
  static constexpr int OFF_HS = OFF_W + RU_STAGES * A_BYTES;
  static constexpr int OFF_X = OFF_HS + TILE_BYTES;                    // two residual tiles
  static constexpr int OFF_BAR = OFF_X + 2 * TILE_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;  // +1024: manual alignment
};
static_assert(RuSmem::XS_HALF % 1024 == 0 && RuSmem::OFF_W % 1024 == 0 && RuSmem::OFF_HS % 1024 == 0, "swizzle atoms");
static_assert(RuSmem::TOTAL <= 232448, "res_unit shared memory");
// the MMA warp cannot consume a tile's weight blocks before its xs box has landed: the producer must get to the box
// request without needing a ring slot back (tests/test_resunit_protocol_model.py)
static_assert(RU_LOADA_AT < RU_STAGES, "xs box request within the ring depth");

struct RuParams {
  int L;    // frames (rows)
  int dil;  // dilation of the 7-tap convolution
  const float *bias1, *a2, *ib2;   // conv7 bias, snake2 exp(alpha), 1/(exp(beta)+1e-9)
  const float *bias2, *an, *ibn;   // conv1 bias, next op's Snake
  bf16 *ox, *oxs;                  // x' and snake_next(x'), [L, 128]
};

#ifdef __CUDACC__

#ifdef ACE_RU_TIMING
// probe builds (tools/ru_timing.cu): cycles the MMA warp of CTA 0 waits on each barrier class, and its total
__device__ long long g_ru_wait[8];
__device__ long long g_ru_pwait[8];
__device__ long long g_ru_e2[8];     // epilogue 2, warp 12: wait D2F, wait XF, TMEM + math, x' store, xs' store, total
__device__ long long g_ru_e1[8];     // epilogue 1, warp 4: wait D1F, wait HSE, TMEM + math + smem store, total
#define RU_E2(i) do { if (e2_stamp) { const long long n_ = clock64(); ru_e[i] += n_ - e_prev; e_prev = n_; } } while (0)  // the producer warp: ring slot free, xs buffer free, x buffer free, total
#define RU_WAIT(i, stmt) do { const long long t0_ = clock64(); stmt; if (blockIdx.x == 0) ru_w[i] += clock64() - t0_; } while (0)
#else
#define RU_WAIT(i, stmt) do { stmt; } while (0)
#define RU_E2(i) do { } while (0)
#endif

enum RuBar {
  RU_FULL = 0, RU_EMPTY = RU_STAGES, RU_AF = 2 * RU_STAGES, RU_AE, RU_D1F, RU_D1F1, RU_D1E, RU_D1E1,
  RU_HSF, RU_HSE, RU_D2F, RU_D2F1, RU_D2E, RU_D2E1, RU_XF, RU_XF1, RU_XE, RU_XE1, RU_NBAR
};
static_assert(RU_NBAR <= 30, "barrier block");

__global__ void __launch_bounds__(RU_THREADS, 1)
res_unit_kernel(const __grid_constant__ CUtensorMap tm_xs, const __grid_constant__ CUtensorMap tm_w1,
                const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ CUtensorMap tm_ox, const __grid_constant__ CUtensorMap tm_oxs,
                const RuParams p) {
  using S = RuSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem + S::OFF_W;
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
    tma_prefetch_desc(&tm_ox);
    tma_prefetch_desc(&tm_oxs);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < RU_STAGES; ++i) {
      mbar_init(&bar[RU_FULL + i], 1);
      mbar_init(&bar[RU_EMPTY + i], 1);
    }
    mbar_init(&bar[RU_AF], 1);
    mbar_init(&bar[RU_AE], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar[RU_D1F + i], 1);
      mbar_init(&bar[RU_D1E + i], 8);
      mbar_init(&bar[RU_D2F + i], 1);
      mbar_init(&bar[RU_D2E + i], 8);
      mbar_init(&bar[RU_XF + i], 1);
      mbar_init(&bar[RU_XE + i], 8);
    }
    mbar_init(&bar[RU_HSF], 8);
    mbar_init(&bar[RU_HSE], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // every CTA runs the same number of tile iterations; tiles past the end read zero-filled rows and store nothing
  const int num_tiles = (p.L + 127) / 128;
  const int my_tiles = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int xs_rows = 128 + 6 * p.dil;  // rows of the xs box (= box height of tm_xs)

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    const bool elected = elect_one();
    pdl_wait();
#ifdef ACE_RU_TIMING
    long long ru_w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long ru_t0 = clock64();
#endif
    auto load_x = [&](int j) {  // residual rows of this CTA's j-th tile (two buffers: epilogue 2 is what bounds the
      const int buf = j & 1;    // kernel, so tile j's rows are fetched while it still works on tile j-1's)
      RU_WAIT(2, mbar_wait(&bar[RU_XE + buf], (uint32_t)(((j >> 1) & 1) ^ 1)));
      if (elected) {
        const int m0 = ((int)blockIdx.x + j * (int)gridDim.x) * 128;
        uint8_t* dst = sX + buf * S::TILE_BYTES;
        mbar_arrive_expect_tx(&bar[RU_XF + buf], S::TILE_BYTES);
        tma_load_2d(dst, &tm_x, &bar[RU_XF + buf], 0, m0);
        tma_load_2d(dst + S::A_BYTES, &tm_x, &bar[RU_XF + buf], 64, m0);
      }
    };
    auto load_a = [&](int j) {  // xs box of this CTA's j-th tile: rows [m0 - 3 dil, m0 + 128 + 3 dil), both halves;
      RU_WAIT(1, mbar_wait(&bar[RU_AE], (uint32_t)((j & 1) ^ 1)));  // ONE box: free once tile j-1's conv7 has retired
      if (elected) {
        const int m0 = ((int)blockIdx.x + j * (int)gridDim.x) * 128;
        mbar_arrive_expect_tx(&bar[RU_AF], (uint32_t)(2 * xs_rows * 128));
        tma_load_2d(smem, &tm_xs, &bar[RU_AF], 0, m0 - 3 * p.dil);
        tma_load_2d(smem + S::XS_HALF, &tm_xs, &bar[RU_AF], 64, m0 - 3 * p.dil);
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    auto load_w = [&](const CUtensorMap* tm, int col) {  // one 16 KB weight block [128 out x 64 in] into the ring
      RU_WAIT(0, mbar_wait(&bar[RU_EMPTY + stage], phase ^ 1));
      if (elected) {
        mbar_arrive_expect_tx(&bar[RU_FULL + stage], S::A_BYTES);
        tma_load_2d(sW + stage * S::A_BYTES, tm, &bar[RU_FULL + stage], col, 0);
      }
      if (++stage == RU_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      for (int kb = 0; kb < RU_KB; ++kb) {
        load_w(&tm_w1, (kb >> 1) * RU_C + (kb & 1) * 64);
        // this tile's xs box, behind its first weight blocks (which the ring could take while the previous tile's
        // conv7 was still running): the tensor pipe idles for the box's round trip once per tile, which is free —
        // it is busy 4.1k of the >= 11k cycles epilogue 2 needs per tile
        if (kb == RU_LOADA_AT) load_a(it);
        // the previous tile's residual rows are needed only after its conv1, which the MMA warp issues
        // at K block RU_MMA2_AT of this tile; by now epilogue 2 of the tile before that has long
        // released the buffer, so this wait never stalls the loads that feed the tensor core
        if (kb == RU_LOADX_AT && it > 0) load_x(it - 1);
        if (kb == RU_MMA2_AT && it > 0) {  // W2 for conv1 of the previous tile, in the order the MMA warp consumes
          load_w(&tm_w2, 0);
          load_w(&tm_w2, 64);
        }
      }
    }
    if (my_tiles > 0) {
      load_w(&tm_w2, 0);
      load_w(&tm_w2, 64);
      load_x(my_tiles - 1);
    }
#ifdef ACE_RU_TIMING
    if (blockIdx.x == 0 && elected) {
      for (int i = 0; i < 3; ++i) g_ru_pwait[i] = ru_w[i];
      g_ru_pwait[3] = clock64() - ru_t0;
    }
#endif
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_umma_idesc_bf16(128, RU_C);
    const bool elected = elect_one();
    const uint32_t xs_lo = (smem_u32(smem) >> 4) & 0x3FFFu, w_lo = (smem_u32(sW) >> 4) & 0x3FFFu;
    const uint32_t hs_lo = (smem_u32(sHS) >> 4) & 0x3FFFu;
    int stage = 0;
    uint32_t phase = 0;
    auto next_stage = [&]() {
      if (++stage == RU_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    };
#ifdef ACE_RU_TIMING
    long long ru_w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long ru_t0 = clock64();
#endif
    auto mma2 = [&](int j) {  // x'_j accumulators: D2[j & 1] = hs_j . W2^T   (K = 128 = 2 ring blocks x 4 slices)
      const int b = j & 1;
      RU_WAIT(3, mbar_wait(&bar[RU_HSF], (uint32_t)(j & 1)));
      RU_WAIT(4, mbar_wait(&bar[RU_D2E + b], (uint32_t)(((j >> 1) & 1) ^ 1)));
      const uint32_t d = tmem_base + 256u + (uint32_t)(b * RU_C);
      for (int half = 0; half < 2; ++half) {
        RU_WAIT(0, mbar_wait(&bar[RU_FULL + stage], phase));
        tcgen05_fence_after();
        if (elected) {
          const uint32_t b_lo = w_lo + (uint32_t)stage * (S::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_lo(d, hs_lo + (uint32_t)(half * (S::A_BYTES >> 4) + 2 * k), b_lo + 2 * k, idesc,
                            (half | k) != 0 ? 1u : 0u);
          umma_commit(&bar[RU_EMPTY + stage]);
          if (half == 1) {
            umma_commit(&bar[RU_HSE]);
            umma_commit(&bar[RU_D2F + b]);
          }
        }
        next_stage();
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      RU_WAIT(2, mbar_wait(&bar[RU_D1E + b], (uint32_t)(((it >> 1) & 1) ^ 1)));
      RU_WAIT(1, mbar_wait(&bar[RU_AF], (uint32_t)(it & 1)));
      tcgen05_fence_after();
      const uint32_t d1 = tmem_base + (uint32_t)(b * RU_C);
      const uint32_t box_lo = xs_lo;
      for (int kb = 0; kb < RU_KB; ++kb) {
        RU_WAIT(0, mbar_wait(&bar[RU_FULL + stage], phase));
        tcgen05_fence_after();
        if (elected) {
          // tap kb >> 1 reads box rows [tap * dil, tap * dil + 128): + tap * dil * 128 B on the start address
          const uint32_t a_lo = box_lo + (uint32_t)((kb & 1) * (S::XS_HALF >> 4) + (kb >> 1) * p.dil * 8);
          const uint32_t b_lo = w_lo + (uint32_t)stage * (S::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo(d1, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&bar[RU_EMPTY + stage]);
          if (kb == RU_KB - 1) {
            umma_commit(&bar[RU_D1F + b]);
            umma_commit(&bar[RU_AE]);
          }
        }
        next_stage();
        if (kb == RU_MMA2_AT && it > 0) mma2(it - 1);
      }
    }
    if (my_tiles > 0) mma2(my_tiles - 1);
#ifdef ACE_RU_TIMING
    if (blockIdx.x == 0 && elected) {
      for (int i = 0; i < 5; ++i) g_ru_wait[i] = ru_w[i];
      g_ru_wait[5] = clock64() - ru_t0;
      g_ru_wait[6] = my_tiles;
    }
#endif
  } else if (warp >= 4 && warp < 12) {
    // ---------------- epilogue 1: D1 -> snake2 -> hs tile in smem (A operand of conv1) ----------------
    const int quarter = (warp - 4) & 3, r = quarter * 32 + lane;
    const int c_lo = ((warp - 4) >> 2) * 64;  // this warp's 64 channels
    uint8_t* rowp = sHS + r * 128;
#ifdef ACE_RU_TIMING
    const bool e2_stamp = blockIdx.x == 0 && warp == 4;
    long long ru_e[8] = {0, 0, 0, 0, 0, 0, 0, 0}, e_prev = clock64();
    const long long e_t0 = e_prev;
#endif
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      mbar_wait(&bar[RU_D1F + b], (uint32_t)((it >> 1) & 1));
      RU_E2(0);
      mbar_wait(&bar[RU_HSE], (uint32_t)((it & 1) ^ 1));  // conv1 of the previous tile has read hs
      RU_E2(1);
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
      RU_E2(2);
    }
#ifdef ACE_RU_TIMING
    if (e2_stamp && lane == 0) {
      for (int i = 0; i < 3; ++i) g_ru_e1[i] = ru_e[i];
      g_ru_e1[3] = clock64() - e_t0;
    }
#endif
  } else if (warp >= 12) {
    // ---------------- epilogue 2: D2 + b2 + x -> x', snake_next(x') ----------------
    // Each warp owns a [32 rows x 64 channels] slab of the residual tile in sX (its lane quarter's rows of its
    // channel half): it turns the slab into x' IN PLACE, hands it to a TMA store, then overwrites it with
    // snake_next(x') and hands that to a second one.  Storing thread-per-row straight from registers made every
    // 16-byte store of a warp hit 32 different lines: ncu had the LSU data pipe at 63 % and lg_throttle stalls on
    // this kernel, 92 M store sectors for 2.9 M requests.
    const int quarter = (warp - 12) & 3, r = quarter * 32 + lane;
    const int c_lo = ((warp - 12) >> 2) * 64;  // this warp's 64 channels
    uint8_t* const slab0 = sX + (c_lo >> 6) * S::A_BYTES + quarter * 32 * 128;  // same XOR-by-row swizzle as TMA's
#ifdef ACE_RU_TIMING
    const bool e2_stamp = blockIdx.x == 0 && warp == 12;
    long long ru_e[8] = {0, 0, 0, 0, 0, 0, 0, 0}, e_prev = clock64();
    const long long e_t0 = e_prev;
#endif
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const long row0 = ((long)blockIdx.x + (long)it * gridDim.x) * 128 + quarter * 32;
      const WarpStage slab{slab0 + b * S::TILE_BYTES};
      mbar_wait(&bar[RU_D2F + b], (uint32_t)((it >> 1) & 1));
      RU_E2(0);
      mbar_wait(&bar[RU_XF + b], (uint32_t)((it >> 1) & 1));
      RU_E2(1);
      tcgen05_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + (uint32_t)(b * RU_C);
      uint32_t xs_packed[32];  // snake_next(x') of this thread's 64 channels, held until x' has left the slab
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = c_lo + half * 32;
        float v[32];
        tmem_ld_32x32(taddr + (uint32_t)c, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4* xp = slab.at(lane, half * 4 + q);
          const uint4 xq = *xp;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias2 + c) + 2 * q);
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias2 + c) + 2 * q + 1);
          // x' = bf16(bf16(acc + b2) + x): one pair-wise rounding, then a packed bf16 add (exactly rounded)
          uint4 xn;
          xn.x = badd2(pack_bf16x2(v[8 * q + 0] + b0.x, v[8 * q + 1] + b0.y), xq.x);
          xn.y = badd2(pack_bf16x2(v[8 * q + 2] + b0.z, v[8 * q + 3] + b0.w), xq.y);
          xn.z = badd2(pack_bf16x2(v[8 * q + 4] + b1.x, v[8 * q + 5] + b1.y), xq.z);
          xn.w = badd2(pack_bf16x2(v[8 * q + 6] + b1.z, v[8 * q + 7] + b1.w), xq.w);
          *xp = xn;
          unpack_bf16x2(xn.x, v[8 * q + 0], v[8 * q + 1]); unpack_bf16x2(xn.y, v[8 * q + 2], v[8 * q + 3]);
          unpack_bf16x2(xn.z, v[8 * q + 4], v[8 * q + 5]); unpack_bf16x2(xn.w, v[8 * q + 6], v[8 * q + 7]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.an + c) + i);
          const float4 i4 = __ldg(reinterpret_cast<const float4*>(p.ibn + c) + i);
          xs_packed[half * 16 + 2 * i] = pack_bf16x2(snake_f(v[4 * i + 0], a4.x, i4.x), snake_f(v[4 * i + 1], a4.y, i4.y));
          xs_packed[half * 16 + 2 * i + 1] = pack_bf16x2(snake_f(v[4 * i + 2], a4.z, i4.z), snake_f(v[4 * i + 3], a4.w, i4.w));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[RU_D2E + b]);  // the accumulator is in registers / shared memory
      RU_E2(2);
      // the slab ([32 rows x 64 channels], one swizzle-atom-aligned 4 KB piece of the residual tile) leaves by a TMA
      // store of its own — tensor maps with [64 x 32] boxes; rows past L are clipped — issued by one lane of THIS
      // warp: no cross-warp barrier, and the warp does not spend 8 + 8 load/store pairs per output on the copy
      const int row0i = (int)row0;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm_ox, slab.base, c_lo, row0i);
        tma_store_commit();
        tma_store_wait_read_all();  // x' has been read out of the slab
      }
      __syncwarp();
      RU_E2(3);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *slab.at(lane, q) = make_uint4(xs_packed[4 * q], xs_packed[4 * q + 1], xs_packed[4 * q + 2], xs_packed[4 * q + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm_oxs, slab.base, c_lo, row0i);
        tma_store_commit();
        tma_store_wait_read_all();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[RU_XE + b]);
      RU_E2(4);
    }
#ifdef ACE_RU_TIMING
    if (e2_stamp && lane == 0) {
      for (int i = 0; i < 5; ++i) g_ru_e2[i] = ru_e[i];
      g_ru_e2[5] = clock64() - e_t0;
    }
#endif
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// xs, x: [L, 128] bf16 (xs = snake1(x)); w1: [128, 7*128] tap-major; w2: [128, 128]; outputs [L, 128].
inline int launch_res_unit_fused(const bf16* xs, const bf16* x, const bf16* w1, const bf16* w2, const RuParams& p,
                                 cudaStream_t stream) {
  if (p.L <= 0) return ACE_OK;
  CUtensorMap tm_xs, tm_w1, tm_w2, tm_x, tm_ox, tm_oxs;
  if (p.dil < 1 || p.dil > RU_MAX_DIL) {
    set_error("res_unit: dilation %d outside [1, %d]", p.dil, RU_MAX_DIL);
    return ACE_ERR_INVALID;
  }
  ACE_PROPAGATE(encode_tmap_2d(&tm_xs, xs, RU_C, (uint64_t)p.L, RU_C * sizeof(bf16), 128 + 6 * p.dil));
  ACE_PROPAGATE(encode_tmap_2d(&tm_w1, w1, 7 * RU_C, RU_C, 7 * RU_C * sizeof(bf16), 128));
  ACE_PROPAGATE(encode_tmap_2d(&tm_w2, w2, RU_C, RU_C, RU_C * sizeof(bf16), 128));
  ACE_PROPAGATE(encode_tmap_2d(&tm_x, x, RU_C, (uint64_t)p.L, RU_C * sizeof(bf16), 128));
  ACE_PROPAGATE(encode_tmap_2d(&tm_ox, p.ox, RU_C, (uint64_t)p.L, RU_C * sizeof(bf16), 32));   // one warp's slab
  ACE_PROPAGATE(encode_tmap_2d(&tm_oxs, p.oxs, RU_C, (uint64_t)p.L, RU_C * sizeof(bf16), 32));
  static bool attr_set = false;
  if (!attr_set) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(res_unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RuSmem::TOTAL));
    attr_set = true;
  }
  const int tiles = (p.L + 127) / 128;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  prof_tag_gemm(p.L, RU_C, 8 * RU_C);
  prof_begin(PROF_GEMM, 2.0 * p.L * RU_C * 8.0 * RU_C, 4.0 * p.L * RU_C * 2.0, stream);
  ACE_CUDA_CHECK(launch_kernel(res_unit_kernel, dim3(grid), dim3(RU_THREADS), (size_t)RuSmem::TOTAL, stream, tm_xs,
                               tm_w1, tm_w2, tm_x, tm_ox, tm_oxs, p));
  prof_end(stream);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

#endif  // __CUDACC__

}  // namespace ace
