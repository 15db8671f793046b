// attention_tc.cu — tcgen05 / TMEM flash attention for the DiT (head_dim 128, GQA, bf16).
//
// One CTA = 128 query rows of one (batch, head); keys/values are visited in blocks of 64.
//   warp 0      TMA producer: Q once, then K_j / V_j tiles (3-D tensor maps over the token-major
//               buffers: [columns, tokens, batch], so rows past a batch item's end read as zero)
//   warp 1      single-thread MMA issuer:
//                 S_j = Q · K_j^T    (A = Q K-major, B = K_j K-major)      -> TMEM S[j&1]  (128x64)
//                 O  += P_j · V_j    (A = P_j from smem, B = V_j MN-major) -> TMEM O       (128x128)
//               S_{j+1} is issued before PV_j so it overlaps softmax_j.
//   warp 2      TMEM allocator (256 columns: S0 S1 O)
//   warps 4-11  softmax, TWO threads per query row (each owns 32 of the block's 64 keys; the row
//               maximum is combined through smem + a 64-thread named barrier — the single-warp-
//               per-scheduler version was ALU-latency bound in the round-1 ncu capture, profiles/r1_v3_ncu_full_summary.txt):
//               tcgen05.ld S -> scale (+ band / tail mask only on boundary blocks) -> exp2 ->
//               bf16 P written back to TENSOR MEMORY over the S columns it came from (tcgen05.st) and read
//               by the PV MMA as its A operand (tcgen05.mma with A in TMEM): no shared-memory store, no
//               proxy fence; in-situ A/B on B200: DiT step -3 % at 60 s, -6 % at 240 s.  (The NSW = 8
//               variant and ACE_ATTN_PTMEM=0 keep the swizzled-smem P buffer.)  O stays in TMEM across the
//               whole KV loop; it is rescaled (tcgen05.ld/st) only when a row maximum grows by more
//               than 2^8 (lazy rescale: P and the row sum keep using the stale maximum, which is
//               exact because the final 1/l normalisation uses the same reference).  The normalised output
//               tile is written as bf16 into the (by then idle) Q tile and leaves by two TMA stores.
// Semantics = sdpa with the reference's masks (modeling_acestep_v15_turbo.py:286-368, 1405-1437):
// full, +-window band (|i-j| <= W), or cross attention; softmax in fp32, P rounded to bf16.
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"

namespace ace {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 64;
// threads = 4 control warps + NSW softmax warps
constexpr int Q_HALF = 128 * 64 * 2;   // [128 x 64] bf16 box (16 KB)
constexpr int KV_HALF = 64 * 64 * 2;   // [64 x 64] bf16 box (8 KB)
constexpr int KV_TILE = 2 * KV_HALF;   // [64 x 128]
constexpr int P_BYTES = 128 * 64 * 2;  // [128 q x 64 keys]
// smem: Q (32 KB) | K0 K1 (32 KB) | V0 V1 (32 KB) | P (16 KB) | barriers
constexpr int OFF_K = 2 * Q_HALF;
constexpr int OFF_V = OFF_K + 2 * KV_TILE;
constexpr int OFF_P = OFF_V + 2 * KV_TILE;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int OFF_XCH = OFF_BAR + 256;  // NSW == 8 only: float [3][2][128] half-row max (x2 parity) / sum exchange
template <int NSW>
constexpr int smem_bytes() { return NSW == 8 ? OFF_XCH + 3 * 2 * 128 * 4 : OFF_BAR + 256; }
constexpr float RESCALE_TAU = 8.0f;  // log2 domain

enum Bar {
  Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11,
  P_FULL = 13, P_EMPTY = 15, NBAR = 16  // P_FULL: one barrier (P in smem) or one per S buffer (P in TMEM)
};

// MN-major SWIZZLE_128B operand: atoms of 64 (MN) x 8 (K); SBO = 1024 B between 8-row K groups,
// LBO = bytes between consecutive 64-wide MN atoms (= one [64 keys x 64 d] box here).
__device__ __forceinline__ uint64_t make_umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

#ifdef ACE_ATTN_TIMING
// probe builds only (tools/attn_timing.cu): cycles CTA 0 / warp 4 / lane 0 spends per phase of the KV loop
__device__ long long g_attn_cycles[8];
__device__ unsigned long long g_attn_stamp[8];  // globaltimer (ns) of CTA (0,0,0): entry, setup done, S_0 seen, loop end, last PV, stores done, exit
__device__ __forceinline__ unsigned long long att_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ATT_STAMP(i, cond) do { if ((cond) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_attn_stamp[i] = att_gtimer(); } while (0)
#define ATT_T(i) do { if (stamp) { const long long now_ = clock64(); g_attn_cycles[i] += now_ - tprev; tprev = now_; } } while (0)
#else
#define ATT_T(i) do { } while (0)
#define ATT_STAMP(i, cond) do { } while (0)
#endif

__device__ __forceinline__ void tmem_st_32x32_u32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// PT = true (NSW == 4 only): P_j is written back into tensor memory over the first 32 columns of the S buffer
// it was computed from (64 bf16 keys = 32 packed words per row) and P.V reads its A operand from there: no
// shared-memory store, no generic->async proxy fence, and the P.V MMAs read half as much shared memory.
// Barrier rules of this variant (each one was a race found by tools/attn_stress.py): P_FULL is one mbarrier per S
// buffer; every softmax warp waits for PV_{j-1} in every block; S_j waits for PV_{j-2} before overwriting P_{j-2}.
template <int NSW, bool PT>
__global__ void __launch_bounds__(128 + 32 * NSW, NSW == 4 ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                    const AttnParams p) {
  // No static __shared__ in this kernel, so the dynamic window starts 1024-byte aligned (needed by
  // SWIZZLE_128B); checked rather than padded so that two CTAs fit on one SM.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint8_t* sP = smem + OFF_P;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / p.group;
  ATT_STAMP(0, threadIdx.x == 0);

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bar[Q_FULL], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar[K_FULL + i], 1);
      mbar_init(&bar[K_EMPTY + i], 1);
      mbar_init(&bar[V_FULL + i], 1);
      mbar_init(&bar[V_EMPTY + i], 1);
      mbar_init(&bar[S_FULL + i], 1);
      mbar_init(&bar[S_EMPTY + i], NSW);
    }
    mbar_init(&bar[P_FULL], NSW);
    mbar_init(&bar[P_FULL + 1], NSW);
    mbar_init(&bar[P_EMPTY], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128u;
  pdl_wait();
  ATT_STAMP(1, threadIdx.x == 0);

  // keys visited by this query tile; `skv` = this sample's valid (non-padding) keys
  int skv = p.Skv;
  if (p.kv_len != nullptr) {
    skv = p.kv_len[b];
    skv = skv < 1 ? 1 : (skv > p.Skv ? p.Skv : skv);
  }
  int j_lo = 0, j_hi = skv;
  if (p.window >= 0) {
    j_lo = q0 - p.window;
    if (j_lo < 0) j_lo = 0;
    j_hi = q0 + BQ - 1 + p.window + 1;
    if (j_hi > skv) j_hi = skv;
    if (j_lo >= j_hi) j_lo = j_hi - 1;  // query tile entirely past the valid keys: one (fully masked-in-band) block
  }
  const int nblk = (j_hi - j_lo + BKV - 1) / BKV;

  // Both issue loops run warp-uniformly and predicate only the asynchronous instructions on one
  // elected lane, so descriptors and barrier addresses stay in uniform registers (see elect_one).
  if (warp == 0) {
    // ---------------- TMA producer ----------------
    const bool elected = elect_one();
    if (elected) {
      mbar_arrive_expect_tx(&bar[Q_FULL], 2 * Q_HALF);
      tma_load_3d(sQ, &tm_q, &bar[Q_FULL], h * HD, q0, b);
      tma_load_3d(sQ + Q_HALF, &tm_q, &bar[Q_FULL], h * HD + 64, q0, b);
    }
    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int row = j_lo + j * BKV;
      mbar_wait(&bar[K_EMPTY + s], ph ^ 1);
      if (elected) {
        mbar_arrive_expect_tx(&bar[K_FULL + s], KV_TILE);
        tma_load_3d(sK + s * KV_TILE, &tm_k, &bar[K_FULL + s], hk * HD, row, b);
        tma_load_3d(sK + s * KV_TILE + KV_HALF, &tm_k, &bar[K_FULL + s], hk * HD + 64, row, b);
      }
      mbar_wait(&bar[V_EMPTY + s], ph ^ 1);
      if (elected) {
        mbar_arrive_expect_tx(&bar[V_FULL + s], KV_TILE);
        tma_load_3d(sV + s * KV_TILE, &tm_v, &bar[V_FULL + s], hk * HD, row, b);
        tma_load_3d(sV + s * KV_TILE + KV_HALF, &tm_v, &bar[V_FULL + s], hk * HD + 64, row, b);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_qk = make_umma_idesc_bf16(128, BKV);               // A, B K-major
    constexpr uint32_t idesc_pv = make_umma_idesc_bf16(128, 128) | (1u << 16);  // B MN-major
    const bool elected = elect_one();
    const uint32_t q_lo = (smem_u32(sQ) >> 4) & 0x3FFFu, k_lo = (smem_u32(sK) >> 4) & 0x3FFFu;
    const uint32_t p_lo = (smem_u32(sP) >> 4) & 0x3FFFu;
    // MN-major V: LBO (bytes between the two 64-wide head_dim atoms) sits in bits [16,30) of the low word
    const uint32_t v_lo = ((smem_u32(sV) >> 4) & 0x3FFFu) | ((uint32_t)(KV_HALF >> 4) << 16);
    mbar_wait(&bar[Q_FULL], 0);
    auto issue_s = [&](int j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&bar[K_FULL + s], ph);
      mbar_wait(&bar[S_EMPTY + s], ph ^ 1);
      // P_{j-2} lives in this S buffer (PT): S_j overwrites it, so wait until PV_{j-2} has retired rather than
      // rely on issue order alone (PV_{j-2} was issued a whole softmax earlier: the wait is almost always free).
      if (PT && j >= 2) mbar_wait(&bar[P_EMPTY], (uint32_t)(j & 1));
      tcgen05_fence_after();
      if (elected) {
        const uint32_t d = tmem_base + (uint32_t)(s * BKV);
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // head_dim halves of 64
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss_lo(d, q_lo + (uint32_t)(half * (Q_HALF >> 4) + 2 * k),
                            k_lo + (uint32_t)(s * (KV_TILE >> 4) + half * (KV_HALF >> 4) + 2 * k), idesc_qk,
                            (half | k) != 0 ? 1u : 0u);
        }
        umma_commit(&bar[K_EMPTY + s]);
        umma_commit(&bar[S_FULL + s]);
      }
    };
    issue_s(0);
    for (int j = 0; j < nblk; ++j) {
      if (j + 1 < nblk) issue_s(j + 1);
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&bar[V_FULL + s], ph);
      // PT: P_j has its own barrier per S buffer.  With a single barrier a softmax warp that runs one block ahead
      // (S_{j+1} is ready before PV_j, and nothing makes it wait for P_EMPTY unless it rescales) would put its
      // P_{j+1} arrival into phase j, completing it while a slower warp's P_j is still unwritten.
      if (PT) mbar_wait(&bar[P_FULL + s], ph);
      else mbar_wait(&bar[P_FULL], (uint32_t)(j & 1));
      tcgen05_fence_after();
      if (elected) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // 16 keys per MMA: P advances 32 B inside its swizzled row (8 TMEM columns), V advances 16 rows (2048 B)
          if (PT) {
            umma_bf16_ts_lo(tmem_o, tmem_base + (uint32_t)(s * BKV + 8 * k),
                            v_lo + (uint32_t)(s * (KV_TILE >> 4) + k * (2048 >> 4)), idesc_pv, (j | k) != 0 ? 1u : 0u);
          } else {
            umma_bf16_ss_lo(tmem_o, p_lo + (uint32_t)(2 * k),
                            v_lo + (uint32_t)(s * (KV_TILE >> 4) + k * (2048 >> 4)), idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&bar[V_EMPTY + s]);
        umma_commit(&bar[P_EMPTY]);
      }
    }
  } else if (warp >= 4) {
    // ---------------- softmax + output ----------------
    // NSW = 4: one thread per query row (64 keys of each block), CTA small enough for 2 per SM.
    // NSW = 8: two threads per row — warps 4-7 take keys [0,32) of each block, warps 8-11 keys
    //          [32,64); the half-row maxima / sums are combined through smem + a 64-thread named
    //          barrier.  Warp w may only touch TMEM lanes 32*(w%4)..+31, so both warps of a pair
    //          share the lane quarter (w-4)&3.
    constexpr int HPT = 8 / NSW;          // 32-key chunks per thread
    constexpr int OCH = 4 / (NSW / 4);    // 32-column chunks of O per thread
    const int quarter = (warp - 4) & 3;
    const int hsel = (warp - 4) >> 2;     // 0 for NSW == 4
    const int r = quarter * 32 + lane;    // row within the tile
    const int qi = q0 + r;                // query index within the batch item
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem + OFF_XCH);  // [2 parity + 1][2 halves][128 rows] (NSW == 8)
    // keys this row may attend to: [k_lo, k_hi)
    int k_lo = 0, k_hi = skv;
    if (p.window >= 0) {
      k_lo = qi - p.window > 0 ? qi - p.window : 0;
      k_hi = qi + p.window + 1 < skv ? qi + p.window + 1 : skv;
    }
    const unsigned span = k_hi > k_lo ? (unsigned)(k_hi - k_lo) : 0u;
    float m_used = -INFINITY, l_run = 0.f;
#ifdef ACE_ATTN_TIMING
    const bool stamp = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
    long long tprev = clock64();
#endif

    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int jb0 = j_lo + j * BKV;
      mbar_wait(&bar[S_FULL + s], ph);
      if (j == 0) ATT_STAMP(2, threadIdx.x == 128);
      ATT_T(0);  // wait for S_j
      tcgen05_fence_after();
      __syncwarp();
      float v[HPT][32];
      {  // both 32-column loads in flight, one wait
        uint32_t raw[HPT][32];
#pragma unroll
        for (int c = 0; c < HPT; ++c)
          tmem_ld_32x32_nowait(tmem_base + lane_base + (uint32_t)(s * BKV + (hsel * HPT + c) * 32), raw[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < HPT; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c][i] = __uint_as_float(raw[c][i]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[S_EMPTY + s]);  // this warp's slice of S_j is in registers
      ATT_T(1);  // TMEM load

      // boundary blocks only: tail of the key range and the +-window band (tile-uniform test)
      const bool interior = (jb0 + BKV <= skv) &&
                            (p.window < 0 || ((q0 + BQ - 1) - jb0 <= p.window && (jb0 + BKV - 1) - q0 <= p.window));
      if (!interior) {
#pragma unroll
        for (int c = 0; c < HPT; ++c) {
          const unsigned lo = (unsigned)(k_lo - (jb0 + (hsel * HPT + c) * 32));
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if ((unsigned)i - lo >= span) v[c][i] = -INFINITY;  // key outside [k_lo, k_hi)
        }
      }
      // row maximum: four independent chains (a single fmaxf chain over 64 values is 64 dependent ops)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < HPT; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx4[0] = fmaxf(mx4[0], v[c][i]);
          mx4[1] = fmaxf(mx4[1], v[c][i + 1]);
          mx4[2] = fmaxf(mx4[2], v[c][i + 2]);
          mx4[3] = fmaxf(mx4[3], v[c][i + 3]);
        }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if (NSW == 8) {  // combine the two half-row maxima (raw score domain; scale > 0 commutes with max)
        xch[(s * 2 + hsel) * 128 + r] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        mx = fmaxf(mx, xch[(s * 2 + (hsel ^ 1)) * 128 + r]);
      }
      mx *= p.scale_log2;

      // lazy rescale: keep the stale reference maximum unless some row of this warp outgrew it
      // (both warps of a pair see identical per-row values, so they take the same branch)
      const float m_new = fmaxf(m_used, mx);
      const bool grow = (m_new - m_used > RESCALE_TAU) || (m_used == -INFINITY && m_new != -INFINITY);
      const bool warp_grow = __any_sync(0xffffffffu, grow);
      float factor = 1.f;
      if (warp_grow) {
        factor = (m_used == -INFINITY) ? 0.f : exp2f(m_used - m_new);
        m_used = m_new;
        l_run *= factor;
      }
      const float neg_ref = (m_used == -INFINITY) ? 0.f : -m_used;
      ATT_T(2);  // mask + max + rescale decision

      // P_j is computed into registers BEFORE waiting for the P buffer, so the exponentials of this
      // block overlap the PV MMA of the previous one (which still reads the buffer)
      uint32_t w[HPT][16];
      float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < HPT; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(fmaf(v[c][i], p.scale_log2, neg_ref));
          const float p1 = ex2_approx(fmaf(v[c][i + 1], p.scale_log2, neg_ref));
          rs[i & 2] += p0;
          rs[(i & 2) + 1] += p1;
          w[c][i >> 1] = pack_bf16x2(p0, p1);
        }
      }
      l_run += (rs[0] + rs[1]) + (rs[2] + rs[3]);

      ATT_T(3);  // exponentials
      if (PT) {
        // P_j goes into the S buffer just read, which nothing else uses until S_{j+2}.  Every warp still waits
        // for PV_{j-1} in every block, not only when it rescales O: the parity waits on P_EMPTY are only
        // unambiguous while no waiter is more than one phase behind.  A warp that skipped this wait could finish
        // its last block while PV_{nblk-2} was still pending, see the (nblk-1) parity of the final wait as
        // already complete, and read O two P.V products short (12 % of launches in tools/attn_stress.py).
        // PV_{j-1} was issued a whole softmax block earlier, so the wait is almost always free.
        if (j > 0) mbar_wait(&bar[P_EMPTY], (uint32_t)((j & 1) ^ 1));
        if (warp_grow && j > 0) {
          tcgen05_fence_after();
#pragma unroll 1
          for (int c = 0; c < OCH; ++c) {
            float o[32];
            tmem_ld_32x32(tmem_o + lane_base + (uint32_t)(hsel * 64 + c * 32), o);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= factor;
            tmem_st_32x32(tmem_o + lane_base + (uint32_t)(hsel * 64 + c * 32), o);
          }
        }
        ATT_T(4);
        uint32_t pw[32];
#pragma unroll
        for (int c = 0; c < HPT; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i) pw[(c * 16 + i) & 31] = w[c][i];
        tmem_st_32x32_u32(tmem_base + lane_base + (uint32_t)(s * BKV), pw);
      } else {
      // P buffer (and O) are free once PV_{j-1} has retired
      mbar_wait(&bar[P_EMPTY], (uint32_t)((j & 1) ^ 1));
      ATT_T(4);  // wait for PV_{j-1}
      if (warp_grow && j > 0) {
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < OCH; ++c) {  // this thread's share of the 128 output columns
          float o[32];
          tmem_ld_32x32(tmem_o + lane_base + (uint32_t)(hsel * 64 + c * 32), o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= factor;
          tmem_st_32x32(tmem_o + lane_base + (uint32_t)(hsel * 64 + c * 32), o);
        }
      }
      uint8_t* rowp = sP + r * 128;
#pragma unroll
      for (int c = 0; c < HPT; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((hsel * HPT + c) * 4 + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(rowp + chunk * 16) =
              make_uint4(w[c][4 * q], w[c][4 * q + 1], w[c][4 * q + 2], w[c][4 * q + 3]);
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[P_FULL + (PT ? s : 0)]);
      ATT_T(5);  // (rescale,) P store, fence, arrive
    }

    if (NSW == 8) {  // row sums: add the partner's partial (both are relative to the same m_used)
      xch[(2 * 2 + hsel) * 128 + r] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      l_run += xch[(2 * 2 + (hsel ^ 1)) * 128 + r];
    }

    // O is complete once the last PV has retired
    ATT_STAMP(3, threadIdx.x == 128);
    mbar_wait(&bar[P_EMPTY], (uint32_t)((nblk - 1) & 1));
    ATT_STAMP(4, threadIdx.x == 128);
    tcgen05_fence_after();
    __syncwarp();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    // The output tile leaves through the Q tile's shared memory (idle: every S MMA has retired) and two TMA stores
    // — thread-per-row 16-byte global stores hit 32 different lines per instruction and took 1.3-2.7 us of a
    // 8-19 us CTA (tools/attn_timing.cu).  Same [128 rows x 64 cols] 128-byte-swizzled boxes TMA loaded Q into;
    // rows past the end of the batch item are clipped by the tensor map.
#pragma unroll 1
    for (int c = 0; c < OCH; ++c) {
      float o[32];
      tmem_ld_32x32(tmem_o + lane_base + (uint32_t)(hsel * 64 + c * 32), o);
      const int col0 = hsel * 64 + c * 32;
      uint8_t* rowp = sQ + (col0 >> 6) * Q_HALF + r * 128;
      const int ch0 = (col0 & 63) >> 3;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(o[8 * q + 0] * inv, o[8 * q + 1] * inv);
        w.y = pack_bf16x2(o[8 * q + 2] * inv, o[8 * q + 3] * inv);
        w.z = pack_bf16x2(o[8 * q + 4] * inv, o[8 * q + 5] * inv);
        w.w = pack_bf16x2(o[8 * q + 6] * inv, o[8 * q + 7] * inv);
        *reinterpret_cast<uint4*>(rowp + (((ch0 + q) ^ (r & 7)) << 4)) = w;
      }
    }
    fence_proxy_async_smem();  // this thread's smem writes -> visible to the TMA store
    asm volatile("bar.sync 6, %0;" ::"n"(32 * NSW) : "memory");
    if (warp == 4 && elect_one()) {
      tma_store_3d(&tm_o, sQ, h * HD, q0, b);
      tma_store_3d(&tm_o, sQ + Q_HALF, h * HD + 64, q0, b);
      tma_store_commit();
      tma_store_wait_read_all();
    }
    ATT_STAMP(5, threadIdx.x == 128);
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
  ATT_STAMP(6, threadIdx.x == 64);
}

}  // namespace

#ifdef ACE_PROBE
static int g_attn_ptmem_override = -1;
void set_attention_p_in_tmem(int mode) { g_attn_ptmem_override = mode < 0 ? -1 : (mode != 0); }
#endif

int make_attn_plan(AttnPlan* plan, const AttnParams& p, int heads, int batch) {
  plan->p = p;
  plan->heads = heads;
  plan->batch = batch;
  const int kvh = heads / p.group;
  ACE_PROPAGATE(encode_tmap_3d(&plan->tm_q, p.q, (uint64_t)heads * HD, (uint64_t)p.Sq, (uint64_t)batch,
                               (uint64_t)p.ldq * 2, (uint64_t)p.Sq * p.ldq * 2, BQ));
  ACE_PROPAGATE(encode_tmap_3d(&plan->tm_k, p.k, (uint64_t)kvh * HD, (uint64_t)p.Skv, (uint64_t)batch,
                               (uint64_t)p.ldk * 2, (uint64_t)p.Skv * p.ldk * 2, BKV));
  ACE_PROPAGATE(encode_tmap_3d(&plan->tm_v, p.v, (uint64_t)kvh * HD, (uint64_t)p.Skv, (uint64_t)batch,
                               (uint64_t)p.ldv * 2, (uint64_t)p.Skv * p.ldv * 2, BKV));
  ACE_PROPAGATE(encode_tmap_3d(&plan->tm_o, p.o, (uint64_t)heads * HD, (uint64_t)p.Sq, (uint64_t)batch,
                               (uint64_t)p.ldo * 2, (uint64_t)p.Sq * p.ldo * 2, BQ));
  return ACE_OK;
}

template <int NSW, bool PT>
static int launch_attention_tc_n(const AttnPlan& plan, dim3 grid, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel<NSW, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem_bytes<NSW>()));
    attr = true;
  }
  ACE_CUDA_CHECK(launch_kernel(attention_tc_kernel<NSW, PT>, grid, dim3(128 + 32 * NSW), (size_t)smem_bytes<NSW>(),
                               stream, plan.tm_q, plan.tm_k, plan.tm_v, plan.tm_o, plan.p));
  return ACE_OK;
}

int launch_attention_tc(const AttnPlan& plan, cudaStream_t stream) {
  const AttnParams& p = plan.p;
  if (p.Sq <= 0 || p.Skv <= 0 || plan.batch <= 0) return ACE_OK;
  dim3 grid(ceil_div(p.Sq, BQ), plan.heads, plan.batch);
  const double keys = p.window >= 0 ? (double)(p.Skv < 2 * p.window + 1 ? p.Skv : 2 * p.window + 1)
                                    : (double)p.Skv;
  const int kvh = plan.heads / p.group;
  prof_begin(PROF_ATTN, 4.0 * p.Sq * keys * HD * plan.heads * plan.batch,
             2.0 * HD * plan.batch * ((double)p.Sq * plan.heads * 2 + (double)p.Skv * kvh * 2), stream);
  // The shipped variant: 4 softmax warps per CTA (two co-resident CTAs per SM overlap each other's softmax and
  // MMA phases), P written back to tensor memory.  It measured faster than the 8-softmax-warp / 1-CTA-per-SM
  // variant at both the 60 s (6.19 vs 6.42 ms per step) and 240 s (21.7 vs 22.7 ms) shapes.  Probe builds keep the
  // alternatives selectable (ACE_ATTN_NSW=8, ACE_ATTN_PTMEM=0 / ace_debug_set_attention_p_in_tmem).
#ifdef ACE_PROBE
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("ACE_ATTN_NSW");
    forced = e ? atoi(e) : 0;
  }
  const int nsw = forced == 8 ? 8 : 4;
  static int ptmem_env = -1;
  if (ptmem_env < 0) {
    const char* e = getenv("ACE_ATTN_PTMEM");
    ptmem_env = (e && e[0] == '0') ? 0 : 1;
  }
  const int ptmem = g_attn_ptmem_override >= 0 ? g_attn_ptmem_override : ptmem_env;
  const int st = nsw == 8 ? launch_attention_tc_n<8, false>(plan, grid, stream)
                          : (ptmem ? launch_attention_tc_n<4, true>(plan, grid, stream)
                                   : launch_attention_tc_n<4, false>(plan, grid, stream));
#else
  const int st = launch_attention_tc_n<4, true>(plan, grid, stream);
#endif
  prof_end(stream);
  return st;
}

}  // namespace ace
