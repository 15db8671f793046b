// gemm.cuh — the one tensor-core kernel of libacestep_b200.
//
// A persistent, warp-specialised tcgen05 GEMM for sm_100a:
//
//     D[m, n] = sum_{tap} sum_{k < Kc}  A[m + shift[tap], k] * B[n, tap*Kc + k]      (fp32 in TMEM)
//
//   * A  : bf16, row-major [rows_A, Kc] with an arbitrary row pitch (K-major operand).
//          Rows outside [0, rows_A) read as zero (TMA out-of-bounds fill), which is what makes
//          the "taps" useful: a 1-D convolution over a channels-last signal is this GEMM with
//          shift[tap] = (tap - centre) * dilation, a stride-s transposed convolution is the
//          2-tap case {0, -1}, a plain nn.Linear is the 1-tap case.
//   * B  : bf16, row-major [N, ntaps*Kc]  (nn.Linear weight layout, K-major operand).
//   * D  : never stored raw — an epilogue functor consumes the fp32 accumulator rows straight
//          out of TMEM (tcgen05.ld, one accumulator row per thread) and writes the fused result.
//
// Pipeline: warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B), warp 1 = single-thread
// tcgen05.mma issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.  STAGES-deep smem ring
// (full/empty mbarriers) and a 2-deep TMEM accumulator ring (tmem_full/tmem_empty) so the
// epilogue of tile i overlaps the main loop of tile i+1.  Grid = min(#tiles, #SMs).
//
// Replaces the cuBLAS calls behind every nn.Linear of the reference DiT layer
// (acestep/models/turbo/modeling_acestep_v15_turbo.py:276-279, Qwen3MLP) and the cuDNN
// convolutions of AutoencoderOobleck (structure: acestep/models/mlx/vae_model.py:62-230).
#pragma once

#include "common.cuh"
#include "epilogues.cuh"  // WarpStage + the epilogue functors

namespace ace {

constexpr int GEMM_BM = 128;      // UMMA M (one TMEM lane per accumulator row)
constexpr int GEMM_BK = 64;       // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GEMM_MAX_TAPS = 8;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM2_THREADS = 384;  // CTA-pair kernel: 4 control warps + 8 epilogue warps

struct GemmShape {
  int M, N;             // output extent
  int kblocks_per_tap;  // Kc / 64
  int ntaps;
  int b_tap_stride;     // elements between taps along B's K axis (= Kc)
  int shift[GEMM_MAX_TAPS];
  // Optional L2 prefetch of the NEXT kernel's weights (read-only, so it may start before the
  // programmatic-dependency wait): each CTA's idle warp 3 issues cp.async.bulk.prefetch.L2 for its
  // slice.  Null unless the owner chains plans (dit.cu: off by default, see the A/B note there).
  const uint8_t* pf_ptr;
  unsigned long long pf_bytes;
  // Split-K (single-CTA kernel, BLOCK_N 128 only; see gemm_tc_kernel): `splits` CTAs share one output
  // tile, each accumulating a K slice; partial fp32 tiles go through `part` ([work][128][128]) and the
  // per-tile {arrive, done} counters in `sync`.  splits <= 1: off.
  int splits;
  float* part;
  unsigned* sync;
  // Tile order of the CTA-pair kernel: 0 = m fastest (consecutive clusters share a B tile; B is streamed from HBM
  // once, A is re-read once per wave), 1 = n fastest (consecutive clusters share an A row block; A is streamed
  // once and B — the smaller operand — stays in L2).  make_gemm_plan sets it to "A is the larger operand".
  int raster_n;
  // probe builds (ACE_RACE_DELAY=1): column group 0 of the every-tile tail epilogue sleeps 30 us per tile, which puts
  // group 1 a whole tile ahead of the loader warp — the schedule under which a hand-back protocol that counts
  // arrivals per TILE instead of per tile-with-boxes deadlocks (tests/test_gpu_kernels.py)
  int race_delay;
};

#ifdef __CUDACC__
__device__ __forceinline__ void gemm_prefetch_next(const GemmShape& shp, int lane) {
  if (shp.pf_ptr == nullptr) return;
  constexpr unsigned long long CH = 8192;
  const unsigned long long per_cta = ((shp.pf_bytes + gridDim.x - 1) / gridDim.x + CH - 1) / CH * CH;
  const unsigned long long lo = (unsigned long long)blockIdx.x * per_cta;
  unsigned long long hi = lo + per_cta;
  if (hi > shp.pf_bytes) hi = shp.pf_bytes;
  for (unsigned long long off = lo + (unsigned long long)lane * CH; off < hi; off += 32ull * CH) {
    const unsigned long long n = hi - off < CH ? hi - off : CH;
    l2_prefetch_bulk(shp.pf_ptr + off, (uint32_t)(n & ~15ull));
  }
}
#endif

// Host-side description of one GEMM problem; tensor maps are encoded once (at bind time) so a
// launch does no host work besides cudaLaunchKernel and the whole step can live in a CUDA graph.
struct GemmPlan {
  CUtensorMap tma_a, tma_b;
  GemmShape shp;
  const bf16* a_ptr;
  long a_ld;
  int a_rows;
  const bf16* b_ptr;
  long b_ld;
  int bn;
};

// Encodes a 2-D bf16 tensor map (dim0 = contiguous columns) with a [64 x box_rows] box and
// 128-byte swizzle.  Implemented in runtime.cu (driver entry point looked up at run time).
int encode_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                   uint64_t row_pitch_bytes, uint32_t box_rows);

// 3-D variant [cols, rows, batch] (attention operands); box = [64 x box_rows x 1].
int encode_tmap_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batches,
                   uint64_t row_pitch_bytes, uint64_t batch_pitch_bytes, uint32_t box_rows);

int make_gemm_plan(GemmPlan* plan, const bf16* a, int a_rows, int kc, long a_ld, const bf16* b,
                   int n, long b_ld, int m, int ntaps, const int* shifts, int bn);

// Split-K for problems with few output tiles (M <= a few hundred rows).  SINGLE-STREAM REQUIREMENT: the split-K tail
// spin-waits across CTAs and relies on all `works <= #SMs` CTAs of the launch being co-resident, which holds when the
// handle's kernels run one after another on one stream of an otherwise idle GPU (the DiT step, ace_enc_forward).
// Two split-K GEMMs issued concurrently from different streams / handles / MPS clients could each hold part of the
// SMs while spinning; the bounded wait (ACE_HANG_GUARD) then traps after ~2 s instead of hanging.  Callers that need
// concurrency across handles must serialise small-M (T < ~20 s) steps themselves.  `scratch_bytes(works)` of
// device memory hold the partial tiles followed by the per-tile counters (zero-initialised once by the
// caller; the kernel re-arms them).  gemm_plan_enable_splitk decides from the shape (cost model in
// runtime.cu), may switch a pair-tile plan to single-CTA 128-wide tiles, and leaves the plan untouched
// when splitting would not pay or the scratch is too small.
constexpr int GEMM_SPLITK_MAX_WORKS = 160;
inline size_t gemm_splitk_scratch_bytes() {
  return (size_t)GEMM_SPLITK_MAX_WORKS * GEMM_BM * 128 * sizeof(float) + (size_t)GEMM_SPLITK_MAX_WORKS * 2 * sizeof(unsigned);
}
int gemm_plan_enable_splitk(GemmPlan* plan, void* scratch, size_t scratch_bytes);

// Debug switch (tests only): route launches through the scalar reference kernels below so the
// epilogues and the surrounding pipeline can be validated independently of the tcgen05 main loop.
#ifdef ACE_PROBE
void set_gemm_debug_reference(bool on);
bool gemm_debug_reference();
float* gemm_debug_scratch(size_t elems);  // grows a device scratch buffer; nullptr on failure
#else
constexpr bool gemm_debug_reference() { return false; }
#endif
int num_sms();

#ifdef __CUDACC__

// Accumulator sources: the epilogues are written once against this tiny interface.
struct AccTmem {
  uint32_t taddr;  // lane-quarter base | first column of this tile's accumulator
  __device__ __forceinline__ void load32(int col, float (&v)[32]) const {
    tmem_ld_32x32(taddr + (uint32_t)col, v);
  }
  // issue only; the caller runs tmem_ld_wait() once after several of these
  __device__ __forceinline__ void load32_nowait(int col, uint32_t (&r)[32]) const {
    tmem_ld_32x32_nowait(taddr + (uint32_t)col, r);
  }
};
struct AccSmem {
  const float* p;  // &reduced[row_in_slice * BN + first column of this warp]
  __device__ __forceinline__ void load32(int col, float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 q = *reinterpret_cast<const float4*>(p + col + 4 * i);
      v[4 * i + 0] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  }
};
struct AccGlobal {
  const float* p;  // &scratch[row * ld + n0]
  __device__ __forceinline__ void load32(int col, float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = p[col + i];
  }
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;                 // (2 * STAGES + 4) mbarriers + TMEM slot
  static constexpr int EPI_BYTES = 8 * 4096;            // one 32 x 64 bf16 staging slab per epilogue warp
  static constexpr int OFF_EPI = STAGES * STAGE_BYTES + BAR_BYTES;
  static constexpr int TOTAL = OFF_EPI + EPI_BYTES + 1024;  // +1024: manual align
};

template <int BN, int STAGES, class Epi>
__global__ void __launch_bounds__(GEMM2_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmShape shp, const Epi epi) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  using L = GemmSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 3) gemm_prefetch_next(shp, lane);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], Epi::kHalfTile ? 8 : 4);  // one arrive per active epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Dependents are released only AFTER this CTA owns its tensor memory: a dependent CTA that became resident on
  // this SM first could take the TMEM columns and then sit in griddepcontrol.wait for a grid that can never finish.
  pdl_trigger();
  pdl_wait();  // everything above overlapped the previous kernel's tail

  const int m_tiles = (shp.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (shp.N + BN - 1) / BN;
  const int total_kb = shp.ntaps * shp.kblocks_per_tap;
  // Split-K: work item w = tile * S + s covers K blocks [s * total_kb / S, (s + 1) * total_kb / S).
  // Small-M problems (a 10 s clip is M = 125) have too few output tiles to pull the weights through
  // more than a handful of SMs (16 tiles at N = 2048: 1.7 TB/s of the 6.5 TB/s HBM can reach them);
  // S CTAs per tile stream disjoint K slices instead.  The launcher guarantees works <= #SMs, i.e.
  // every CTA of the grid is resident, which the inter-CTA wait below relies on.
  const int S = shp.splits > 1 ? shp.splits : 1;
  const int num_works = m_tiles * n_tiles * S;

  if (warp == 0) {
    // ---------------- TMA producer (warp-uniform loop, one elected lane issues) ----------------
    const bool elected = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int work = blockIdx.x; work < num_works; work += gridDim.x) {
      const int tile = work / S, sp = work - tile * S;
      const int m0 = (tile % m_tiles) * GEMM_BM;
      const int n0 = (tile / m_tiles) * BN;
      const int kb0 = (int)((long)sp * total_kb / S), kb1 = (int)((long)(sp + 1) * total_kb / S);
      int tap = kb0 / shp.kblocks_per_tap, kk = kb0 - tap * shp.kblocks_per_tap;
      int a_row = m0 + shp.shift[tap];  // refreshed after the loads of a tap's last block (off the issue path)
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elected) {
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          tma_load_2d(sA + stage * L::A_BYTES, &tma_a, &full_bar[stage], kk * GEMM_BK, a_row);
          tma_load_2d(sB + stage * L::B_BYTES, &tma_b, &full_bar[stage],
                      tap * shp.b_tap_stride + kk * GEMM_BK, n0);
        }
        if (++kk == shp.kblocks_per_tap) {
          kk = 0;
          if (++tap < shp.ntaps) a_row = m0 + shp.shift[tap];
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (warp-uniform loop, one elected lane issues) ----------------
    constexpr uint32_t idesc = make_umma_idesc_bf16(GEMM_BM, BN);
    const bool elected = elect_one();
    const uint32_t a_lo0 = (smem_u32(sA) >> 4) & 0x3FFFu, b_lo0 = (smem_u32(sB) >> 4) & 0x3FFFu;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int work = blockIdx.x; work < num_works; work += gridDim.x, ++it) {
      const int sp = work % S;
      const int kb0 = (int)((long)sp * total_kb / S), kb1 = (int)((long)(sp + 1) * total_kb / S);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elected) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (L::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)stage * (L::B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4)
            umma_bf16_ss_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[as]);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4 && (Epi::kHalfTile || warp < 8)) {
    // ---------------- epilogue: TMEM -> registers -> fused op -> HBM ----------------
    // warp w reads TMEM lanes 32*(w%4)..+31.  Epilogues that work on 64-column half tiles use all
    // eight warps (4-7: columns [0,BN/2), 8-11: [BN/2,BN)); head-structured ones use warps 4-7.
    const int quarter = (warp - 4) & 3;
    constexpr int EBN = Epi::kHalfTile ? BN / 2 : BN;
    constexpr int EPI_THREADS = Epi::kHalfTile ? 256 : 128;
    const int sub = Epi::kHalfTile ? ((warp - 4) >> 2) * EBN : 0;
    const int epi_tid = threadIdx.x - 128;
    int it = 0;
    for (int work = blockIdx.x; work < num_works; work += gridDim.x, ++it) {
      const int tile = work / S, sp = work - tile * S;
      const int m0 = (tile % m_tiles) * GEMM_BM;
      const int n0 = (tile / m_tiles) * BN + sub;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const bool live = n0 < shp.N;
      const WarpStage stg{smem + L::OFF_EPI + (warp - 4) * 4096};
      float pre = 0.f;
      if (S == 1 && live) pre = epi.prefetch(m0 + quarter * 32 + lane, n0, shp.M, shp.N, stg);
      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
      __syncwarp();
      AccTmem acc{tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + sub)};
      if (S == 1) {
        if (live) epi.template run<EBN>(acc, m0 + quarter * 32 + lane, n0, shp.M, shp.N, stg, pre);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
        continue;
      }
      // ---- split-K tail (grid == num_works: exactly one work item per CTA) ----
      // (1) this CTA's partial accumulator -> global scratch
      float* mine = shp.part + (size_t)work * (GEMM_BM * BN) + (size_t)(quarter * 32 + lane) * BN + sub;
#pragma unroll 1
      for (int c = 0; c < EBN; c += 32) {
        float v[32];
        acc.load32(c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          __stcg(reinterpret_cast<float4*>(mine + c) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      // (2) wait until all S partials of this tile are in L2.  The CTA barrier orders every warp's stores
      //     before thread 0's gpu-scope fence + arrive (fences are cumulative), and thread 0's acquire
      //     before the other warps' loads on the way back.
      asm volatile("bar.sync 1, %0;" ::"r"(EPI_THREADS) : "memory");
      if (epi_tid == 0) {
        unsigned* arrive = shp.sync + 2 * tile;
        __threadfence();
        atomicAdd(arrive, 1u);
        unsigned seen;
#if ACE_HANG_GUARD
        const long long t0 = clock64();
#endif
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive) : "memory");
#if ACE_HANG_GUARD
          if (seen < (unsigned)S && clock64() - t0 > 4000000000LL) {  // ~2 s: a protocol bug must not wedge the GPU
            printf("ace: split-K wait timeout (block %d tile %d: %u of %d arrived)\n", blockIdx.x, tile, seen, S);
            __trap();
          }
#endif
        } while (seen < (unsigned)S);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(EPI_THREADS) : "memory");
      // (3) rows [sp * R, sp * R + R) of the tile are reduced by this CTA, in split order (deterministic),
      //     into shared memory (the operand stages are idle by now).  All S loads of an element are in
      //     flight together: one L2 round trip per pass.
      const int R = GEMM_BM / S;
      float* red = reinterpret_cast<float*>(smem);
      const float4* tile_part =
          reinterpret_cast<const float4*>(shp.part + (size_t)tile * S * (GEMM_BM * BN) + (size_t)sp * R * BN);
      for (int idx = epi_tid; idx < R * (BN / 4); idx += EPI_THREADS) {
        float4 b[8];
#pragma unroll
        for (int ss = 0; ss < 8; ++ss)
          if (ss < S) b[ss] = __ldcg(tile_part + (size_t)ss * (GEMM_BM * BN / 4) + idx);
        float4 a = b[0];
#pragma unroll
        for (int ss = 1; ss < 8; ++ss)
          if (ss < S) {
            a.x += b[ss].x; a.y += b[ss].y; a.z += b[ss].z; a.w += b[ss].w;
          }
        reinterpret_cast<float4*>(red)[idx] = a;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(EPI_THREADS) : "memory");
      if (epi_tid == 0) {  // last reader of the tile re-arms its counters for the next launch
        unsigned* sy = shp.sync + 2 * tile;
        if (atomicAdd(sy + 1, 1u) == (unsigned)(S - 1)) {
          sy[1] = 0u;
          __threadfence();
          sy[0] = 0u;
        }
      }
      // (4) fused epilogue on the slice: slice row lr of this warp -> global row m0 + sp * R + lr
      if (live && quarter * 32 < R) {
        const int lr = quarter * 32 + lane;
        const int row = m0 + sp * R + lr;
        int m_end = m0 + sp * R + R;
        if (m_end > shp.M) m_end = shp.M;
        const AccSmem racc{red + (size_t)(lr < R ? lr : R - 1) * BN + sub};
        const float spre = epi.prefetch(row, n0, m_end, shp.N, stg);
        epi.template run<EBN>(racc, row, n0, m_end, shp.N, stg, spre);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant: a 2-CTA cluster computes one 256 x 256 output tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256).  CTA r of the pair stages rows [128r, 128r+128) of the
// A tile and rows [128r, 128r+128) of the B tile (half of N); the tensor core reads both halves of
// B from the two CTAs' shared memories, so each CTA pulls 32 KB from L2 per 64-deep K block for
// 2 x 128 x 256 x 64 FLOPs — twice the arithmetic intensity of the 128 x 128 single-CTA tile,
// which measured L2-bandwidth-bound (round-1 A/B measurement; the raw log was not kept — the kernel-level ncu rows of that build are in profiles/r1_v3_ncu_full_summary.txt).  Accumulators: CTA r's TMEM holds its
// 128 rows x 256 columns, double-buffered (2 x 256 = all 512 TMEM columns).
//   barriers: full[s]  (leader only, 2 arrivals: leader expect_tx + peer remote arrive; all four
//                       TMA loads credit the leader's barrier)
//             empty[s], tmem_full[a]  (both CTAs, signalled by multicast tcgen05.commit)
//             tmem_empty[a]           (leader only, 8 arrivals = 4 epilogue warps x 2 CTAs)
// ---------------------------------------------------------------------------------------------
#ifdef ACE_GEMM_TIMING
// probe builds only: per-phase globaltimer stamps of cluster 0 / CTA 0 (ns)
__device__ unsigned long long g_gemm_stamps[16];
__device__ long long g_gemm_cycles[4];  // cluster 0 MMA issuer: clock64 at first data / loop exit, k-blocks issued
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ACE_STAMP(i) do { if (blockIdx.x == 0) g_gemm_stamps[i] = gtimer(); } while (0)
#else
#define ACE_STAMP(i) do { } while (0)
#endif

// ALLTAIL (multi-wave problems of a kTmaTail epilogue): EVERY tile ends in the tail path.  The residual boxes get
// their own shared memory (NBOX x 16 KB after the ring, in place of the per-warp staging slabs; the 256-wide tile
// gives up one ring stage for it), warp 3 loads the next tile's boxes as soon as the epilogue has handed the
// previous ones back (resid_empty), h is updated in place and stored by TMA, then g = h * c is made in place in the
// same box once that store has read it.  The slab path it replaces spent about as long per 256-wide tile as the
// main loop (per-row global loads of the residual, two staged store passes).
template <int BN, int STAGES, class Epi, bool ALLTAIL = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const GemmShape shp, const __grid_constant__ Epi epi) {
  // BN = 256, or 192 for epilogues that accept 64-column sub-tiles: 256 x 192 pair tiles turn the
  // 48-tile N = 2048 problems of the DiT (48 of 74 clusters busy) into 66 tiles (66 of 74).
  static_assert(BN == 256 || (BN == 192 && Epi::kHalfTile), "pair tile width");
  static_assert(!ALLTAIL || Epi::kTmaTail, "ALLTAIL needs a tail-path epilogue");
  constexpr int HALF = BN / 2;       // B rows staged by each CTA of the pair
  using L = GemmSmem<HALF, STAGES>;  // per-CTA stage: 128 A rows + HALF B rows
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int BOX_BYTES = ALLTAIL ? (BN / 64) * L::A_BYTES : 0;  // dedicated residual boxes (ALLTAIL)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * L::A_BYTES;
  uint8_t* boxes = smem + STAGES * L::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES + BOX_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* resid_bar = bars + 2 * STAGES + 6;     // [4] residual box i has landed
  uint64_t* resid_empty = bars + 2 * STAGES + 10;  // [2] ALLTAIL: column group g's boxes may be reloaded

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) ACE_STAMP(0);

  if (warp == 3) gemm_prefetch_next(shp, lane);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);  // 8 epilogue warps x 2 CTAs
    }
    for (int i = 0; i < 4; ++i) mbar_init(&resid_bar[i], 1);  // tail path: one per 64-column residual box
    for (int i = 0; i < 2; ++i) mbar_init(&resid_empty[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // after the TMEM allocation, see gemm_tc_kernel
  pdl_wait();
  if (threadIdx.x == 0) ACE_STAMP(1);

  const int m_tiles = (shp.M + 255) / 256;
  const int n_tiles = (shp.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const bool nfast = shp.raster_n != 0;
  auto tile_m = [&](int t) { return nfast ? t / n_tiles : t % m_tiles; };
  auto tile_n = [&](int t) { return nfast ? t % n_tiles : t / m_tiles; };
  const int total_kb = shp.ntaps * shp.kblocks_per_tap;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  // Tail path (Epi::kTmaTail with its tensor maps set): the last tile of this cluster ends in tail_box() — see the
  // producer's residual loads below and the epilogue's tail branch.
  bool tma_tail = false;
  if constexpr (Epi::kTmaTail) tma_tail = epi.use_tma != 0;
  constexpr int NBOX = BN / 64;
  const int my_tiles = cluster_id < num_tiles ? (num_tiles - 1 - cluster_id) / num_clusters + 1 : 0;
  const int last_tile = cluster_id + (my_tiles - 1) * num_clusters;

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs; warp-uniform loop, one elected lane issues) --------
    const bool elected = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = tile_m(tile) * 256 + (int)rank * 128;
      const int n0 = tile_n(tile) * BN + (int)rank * HALF;
      int tap = 0, kk = 0;
      int a_row = m0 + shp.shift[0];  // refreshed after the loads of a tap's last block (off the issue path)
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elected) {
          if (leader) {
            mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          } else {
            mbar_arrive_remote(&full_bar[stage], 0);
          }
          if (kb == 0 && tile == cluster_id) ACE_STAMP(2);
          tma_load_2d_pair(sA + stage * L::A_BYTES, &tma_a, &full_bar[stage], kk * GEMM_BK, a_row);
          tma_load_2d_pair(sB + stage * L::B_BYTES, &tma_b, &full_bar[stage],
                           tap * shp.b_tap_stride + kk * GEMM_BK, n0);
        }
        if (++kk == shp.kblocks_per_tap) {
          kk = 0;
          if (++tap < shp.ntaps) a_row = m0 + shp.shift[tap];
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if constexpr (Epi::kTmaTail && !ALLTAIL) {
      // Residual boxes of the last tile: they ride the operand ring as NBOX extra "k-blocks" that nobody feeds to
      // the tensor core — box i lands in the A part of the next ring slot as soon as the MMAs that read it have
      // retired (5 k-blocks before the main loop ends), i.e. it is in shared memory before the accumulator is.
      if (tma_tail && my_tiles > 0) {
        const int m0 = tile_m(last_tile) * 256 + (int)rank * 128;
        const int n0 = tile_n(last_tile) * BN;
        for (int i = 0; i < NBOX; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elected && n0 + 64 * i < shp.N) {
            mbar_arrive_expect_tx(&resid_bar[i], L::A_BYTES);
            tma_load_2d(sA + stage * L::A_BYTES, &epi.tm_h, &resid_bar[i], n0 + 64 * i, m0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) -------
      constexpr uint32_t idesc = make_umma_idesc_bf16(256, BN);
      const bool elected = elect_one();
      const uint32_t a_lo0 = (smem_u32(sA) >> 4) & 0x3FFFu, b_lo0 = (smem_u32(sB) >> 4) & 0x3FFFu;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
#ifdef ACE_GEMM_TIMING
      long long c_first = 0, n_kb = 0;
#endif
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (kb == 0 && it == 0) ACE_STAMP(3);
#ifdef ACE_GEMM_TIMING
          if (kb == 0 && it == 0) c_first = clock64();
          ++n_kb;
#endif
          tcgen05_fence_after();
          if (elected) {
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * (L::A_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)stage * (L::B_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4)
              umma_bf16_ss_pair_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[stage], 3);
            if (kb == total_kb - 1) {
              umma_commit_pair(&tfull_bar[as], 3);
              ACE_STAMP(4);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
#ifdef ACE_GEMM_TIMING
      if (blockIdx.x == 0 && elected) {
        g_gemm_cycles[0] = c_first;
        g_gemm_cycles[1] = clock64();
        g_gemm_cycles[2] = n_kb;
      }
#endif
    }
  } else if (warp == 3) {
    if constexpr (ALLTAIL) {
      // ---------------- residual loader (both CTAs): tile i+1's boxes as soon as tile i's are handed back -------
      // Barrier accounting is per box / per column group, never derived from the tile counter: the last column tile
      // of N = 2048 at 192-wide tiles has two boxes, so group 1 has NOTHING to do there.  resid_bar[i] flips only on
      // tiles that have box i, and resid_empty[g] is arrived (and waited for) only for tiles in which group g had
      // boxes — a group that merely passes through a tile must not arrive, or it could complete two phases before
      // this warp has looked at the first one (it then waits for a third that needs this warp's next load: a
      // deadlock that the n-fastest tile order, where such tiles sit in the middle of a CTA's sequence, produced).
      const bool elected = elect_one();
      uint32_t in_use = 0, wait_parity = 0;  // bit g: group g's boxes hold a tile / parity of its next hand-back
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = tile_m(tile) * 256 + (int)rank * 128;
        const int tile_n0 = tile_n(tile) * BN;
        for (int g = 0; g < 2; ++g) {
          if (2 * g >= NBOX || tile_n0 + 128 * g >= shp.N) continue;  // group g has no box in this tile
          if (in_use & (1u << g)) {
            mbar_wait(&resid_empty[g], (wait_parity >> g) & 1u);
            wait_parity ^= 1u << g;
          }
          in_use |= 1u << g;
          if (elected) {
            for (int bi = 2 * g; bi < 2 * g + 2 && bi < NBOX; ++bi) {
              if (tile_n0 + 64 * bi >= shp.N) break;
              mbar_arrive_expect_tx(&resid_bar[bi], L::A_BYTES);
              tma_load_2d(boxes + bi * L::A_BYTES, &epi.tm_h, &resid_bar[bi], tile_n0 + 64 * bi, m0);
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue (both CTAs, each on its own 128 accumulator rows) ----------------
    // warps 4-7 take tile columns [0,128), warps 8-11 columns [128,256); warp w reads TMEM lanes
    // 32*(w%4)..+31.  Operands the epilogue needs from HBM/L2 (residual rows, gates) are prefetched
    // into L1 while the main loop is still running.
    const int quarter = (warp - 4) & 3;
    const int sub = ((warp - 4) >> 2) * 128;
    int it = 0;
    if constexpr (ALLTAIL) {
      const int grp = (warp - 4) >> 2;    // column group: boxes [2 grp, 2 grp + 2) of the tile
      const int r = quarter * 32 + lane;  // row inside this CTA's 128-row tile
      uint32_t box_parity = 0;            // bit i: parity of resid_bar[i]'s next completion
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int m0 = tile_m(tile) * 256 + (int)rank * 128;
        const int tile_n0 = tile_n(tile) * BN;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const auto tpre = epi.tail_prefetch(m0 + r, shp.M, shp.N, tile_n0 + 128 * grp, 128);
        mbar_wait(&tfull_bar[as], aphase);
        if (threadIdx.x == 128) ACE_STAMP(5);
        tcgen05_fence_after();
        __syncwarp();
        AccTmem acc{tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN)};
#pragma unroll 1
        for (int bi = 2 * grp; bi < 2 * grp + 2 && bi < NBOX; ++bi) {
          const int col = tile_n0 + 64 * bi;
          if (col >= shp.N) break;  // group-uniform
          uint8_t* hbox = boxes + bi * L::A_BYTES;
          mbar_wait(&resid_bar[bi], (box_parity >> bi) & 1u);
          box_parity ^= 1u << bi;
          epi.template tail_box<false>(acc, 64 * bi, r, m0 + r, col, shp.M, hbox, nullptr, tpre);
          fence_proxy_async_smem();  // this thread's smem writes -> visible to the TMA store
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (quarter == 0 && elect_one()) {
            tma_store_2d(&epi.tm_h, hbox, col, m0);
            tma_store_commit();
          }
        }
        tcgen05_fence_before();  // the accumulator is free: the next tile's MMAs may overwrite it
#ifdef ACE_PROBE
        if (shp.race_delay && grp == 0) __nanosleep(30000);
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&tempty_bar[as], 0);
        if (epi.no.g != nullptr) {
          if (quarter == 0) tma_store_wait_read_all();  // h's stores have read the boxes
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll 1
          for (int bi = 2 * grp; bi < 2 * grp + 2 && bi < NBOX; ++bi) {
            const int col = tile_n0 + 64 * bi;
            if (col >= shp.N) break;
            epi.tail_g(r, col, boxes + bi * L::A_BYTES, tpre);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (quarter == 0 && elect_one()) {
            for (int bi = 2 * grp; bi < 2 * grp + 2 && bi < NBOX; ++bi) {
              const int col = tile_n0 + 64 * bi;
              if (col >= shp.N) break;
              tma_store_2d(&epi.tm_g, boxes + bi * L::A_BYTES, col, m0);
            }
            tma_store_commit();
          }
        }
        if (quarter == 0 && 2 * grp < NBOX && tile_n0 + 128 * grp < shp.N) {  // this group had boxes in this tile
          tma_store_wait_read_all();  // (only the lane that issued has pending groups)
          __syncwarp();
          if (elect_one()) mbar_arrive(&resid_empty[grp]);
        }
        if (threadIdx.x == 128) ACE_STAMP(6);
      }
      if (quarter == 0) tma_store_wait_read_all();
    } else
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int m0 = tile_m(tile) * 256 + (int)rank * 128;
      const int n0 = tile_n(tile) * BN + sub;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const bool live = n0 < shp.N;
      const WarpStage stg{smem + L::OFF_EPI + (warp - 4) * 4096};
      if constexpr (Epi::kTmaTail && !ALLTAIL) {
        if (tma_tail && tile == last_tile) {
          // ---- tail path: residual already in shared memory (TMA), update in place, h / g leave by TMA ----
          const int grp = (warp - 4) >> 2;                 // column group: boxes [2 grp, 2 grp + 2) of the tile
          const int r = quarter * 32 + lane;               // row inside this CTA's 128-row tile
          const int ring0 = (int)(((long)my_tiles * total_kb) % STAGES);  // ring slot of residual box 0
          const int tile_n0 = tile_n(tile) * BN;
          const auto tpre = epi.tail_prefetch(m0 + r, shp.M, shp.N, tile_n0 + 128 * grp, 128);
          mbar_wait(&tfull_bar[as], aphase);
          if (threadIdx.x == 128) ACE_STAMP(5);
          tcgen05_fence_after();
          __syncwarp();
          AccTmem acc{tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN)};
#pragma unroll 1
          for (int bi = 2 * grp; bi < 2 * grp + 2 && bi < NBOX; ++bi) {
            const int col = tile_n0 + 64 * bi;
            if (col >= shp.N) break;  // group-uniform
            uint8_t* hbox = sA + ((ring0 + bi) % STAGES) * L::A_BYTES;
            uint8_t* gbox = sB + bi * L::A_BYTES;  // the B ring is idle by now (all MMAs retired)
            ACE_TCLK(4 * (bi & 1) + 0);
            mbar_wait(&resid_bar[bi], 0);
            ACE_TCLK(4 * (bi & 1) + 1);
            epi.tail_box(acc, 64 * bi, r, m0 + r, col, shp.M, hbox, gbox, tpre);
            ACE_TCLK(4 * (bi & 1) + 2);
            fence_proxy_async_smem();  // this thread's smem writes -> visible to the TMA store
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            if (quarter == 0 && elect_one()) {
              tma_store_2d(&epi.tm_h, hbox, col, m0);
              if (epi.no.g != nullptr) tma_store_2d(&epi.tm_g, gbox, col, m0);
              tma_store_commit();
            }
            ACE_TCLK(4 * (bi & 1) + 3);
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&tempty_bar[as], 0);
          ACE_TCLK(8);
          if (quarter == 0) tma_store_wait_read_all();  // (only the lane that issued has pending groups)
          ACE_TCLK(9);
          if (threadIdx.x == 128) ACE_STAMP(6);
          continue;
        }
      }
      float pre = 0.f;
      if (live) pre = epi.prefetch(m0 + quarter * 32 + lane, n0, shp.M, shp.N, stg);
      mbar_wait(&tfull_bar[as], aphase);
      if (threadIdx.x == 128) ACE_STAMP(5);
      tcgen05_fence_after();
      __syncwarp();
      if (live) {
        AccTmem acc{tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + sub)};
        if (BN == 256 || sub == 0) {
          epi.template run<128>(acc, m0 + quarter * 32 + lane, n0, shp.M, shp.N, stg, pre);
        } else {
          epi.template run<(BN == 256 ? 128 : BN - 128)>(acc, m0 + quarter * 32 + lane, n0, shp.M, shp.N, stg, pre);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty_bar[as], 0);
      if (threadIdx.x == 128) ACE_STAMP(6);
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
  if (threadIdx.x == 64) ACE_STAMP(7);
}

#ifdef ACE_PROBE
// ---------------------------------------------------------------------------------------------
// Scalar reference path (tests / bring-up only; selected with ace_debug_set_gemm_reference).
// ---------------------------------------------------------------------------------------------
__global__ void gemm_ref_kernel(const bf16* __restrict__ A, long lda, int a_rows,
                                const bf16* __restrict__ B, long ldb, GemmShape shp, int kc,
                                float* __restrict__ out, int ld_out, int m_pad, int n_pad);

template <int BN, class Epi>
__global__ void __launch_bounds__(128)
gemm_ref_epilogue_kernel(const float* __restrict__ scratch, int ld, GemmShape shp, const Epi epi) {
  const int m0 = blockIdx.x * GEMM_BM;
  const int n0 = blockIdx.y * BN;
  __shared__ __align__(16) uint8_t stage_mem[4 * 4096];
  const int row = m0 + threadIdx.x;
  AccGlobal acc{scratch + (size_t)row * ld + n0};
  const WarpStage stg{stage_mem + (threadIdx.x >> 5) * 4096};
  const float pre = epi.prefetch(row, n0, shp.M, shp.N, stg);
  epi.template run<BN>(acc, row, n0, shp.M, shp.N, stg, pre);
}

#endif  // ACE_PROBE

template <int BN, int STAGES, class Epi>
int launch_gemm_bn(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
#ifdef ACE_PROBE
  if (gemm_debug_reference()) {
    const int m_pad = ceil_div(p.shp.M, GEMM_BM) * GEMM_BM;
    const int n_pad = ceil_div(p.shp.N, BN) * BN;
    float* scratch = gemm_debug_scratch((size_t)m_pad * n_pad);
    if (!scratch) {
      set_error("gemm debug scratch allocation failed");
      return ACE_ERR_NOMEM;
    }
    dim3 g(ceil_div(n_pad, 32), ceil_div(m_pad, 8));
    gemm_ref_kernel<<<g, dim3(32, 8), 0, stream>>>(p.a_ptr, p.a_ld, p.a_rows, p.b_ptr, p.b_ld,
                                                   p.shp, p.shp.kblocks_per_tap * GEMM_BK, scratch,
                                                   n_pad, m_pad, n_pad);
    gemm_ref_epilogue_kernel<BN, Epi>
        <<<dim3(m_pad / GEMM_BM, n_pad / BN), 128, 0, stream>>>(scratch, n_pad, p.shp, epi);
    add_launches(2);
    ACE_CUDA_CHECK(cudaGetLastError());
    return ACE_OK;
  }
#endif
  using L = GemmSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, Epi>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int tiles = ceil_div(p.shp.M, GEMM_BM) * ceil_div(p.shp.N, BN);
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (p.shp.splits > 1) {
    if (BN != 128 || tiles * p.shp.splits > num_sms() || p.shp.part == nullptr || p.shp.sync == nullptr) {
      set_error("launch_gemm: invalid split-K plan (bn %d, %d tiles x %d splits)", BN, tiles, p.shp.splits);
      return ACE_ERR_INVALID;
    }
    grid = tiles * p.shp.splits;  // one work item per CTA, all co-resident
  }
  const double ktot = (double)p.shp.ntaps * p.shp.kblocks_per_tap * GEMM_BK;
  prof_tag_gemm(p.shp.M, p.shp.N, (int)ktot);
  prof_begin(PROF_GEMM, 2.0 * p.shp.M * p.shp.N * ktot,
             2.0 * ((double)p.shp.M * p.shp.kblocks_per_tap * GEMM_BK + (double)p.shp.N * ktot +
                    (double)p.shp.M * p.shp.N),
             stream);
  ACE_CUDA_CHECK(launch_kernel(gemm_tc_kernel<BN, STAGES, Epi>, dim3(grid), dim3(GEMM2_THREADS),
                               (size_t)L::TOTAL, stream, p.tma_a, p.tma_b, p.shp, epi));
  prof_end(stream);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

// every-tile tail variant: BN = 256 runs a 5-stage ring (5 x 32 KB + 4 boxes x 16 KB), BN = 192 keeps 6 stages
template <int BN, class Epi>
int launch_gemm_pair_alltail(const GemmPlan& p, const Epi& epi, int grid, cudaStream_t stream) {
  constexpr int STAGES = BN == 256 ? 5 : 6;
  using L = GemmSmem<BN / 2, STAGES>;
  constexpr size_t SMEM = (size_t)STAGES * L::STAGE_BYTES + (BN / 64) * L::A_BYTES + L::BAR_BYTES + 1024;
  static_assert(SMEM <= 232448, "ALLTAIL shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<BN, STAGES, Epi, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_set = true;
  }
  ACE_CUDA_CHECK(launch_kernel(gemm_tc2_kernel<BN, STAGES, Epi, true>, dim3(grid), dim3(GEMM2_THREADS), SMEM, stream,
                               p.tma_a, p.tma_b, p.shp, epi));
  return ACE_OK;
}

template <int BN, int STAGES, class Epi>
int launch_gemm_pair(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  using L = GemmSmem<BN / 2, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<BN, STAGES, Epi>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int tiles = ceil_div(p.shp.M, 256) * ceil_div(p.shp.N, BN);
  int max_clusters = num_sms() / 2;
#ifdef ACE_GEMM_TIMING
  if (const char* e = getenv("ACE_GEMM_MAX_CLUSTERS")) max_clusters = atoi(e);  // probe builds only
#endif
  const int grid = 2 * (tiles < max_clusters ? tiles : max_clusters);
  bool alltail = false;
  if constexpr (Epi::kTmaTail) {
    static const bool off = probe_env("ACE_NO_ALLTAIL") != nullptr;  // probe builds: A/B against the slab path
    static const int kb_max = probe_env("ACE_ALLTAIL_KB") ? atoi(probe_env("ACE_ALLTAIL_KB")) : 1 << 30;
    alltail = epi.use_tma != 0 && tiles > max_clusters && !off &&
              (BN == 192 || p.shp.ntaps * p.shp.kblocks_per_tap <= kb_max);
  }
  const double ktot = (double)p.shp.ntaps * p.shp.kblocks_per_tap * GEMM_BK;
  prof_tag_gemm(p.shp.M, p.shp.N, (int)ktot);
  prof_begin(PROF_GEMM, 2.0 * p.shp.M * p.shp.N * ktot,
             2.0 * ((double)p.shp.M * p.shp.kblocks_per_tap * GEMM_BK + (double)p.shp.N * ktot +
                    (double)p.shp.M * p.shp.N),
             stream);
  if constexpr (Epi::kTmaTail) {
    if (alltail) {
      ACE_PROPAGATE((launch_gemm_pair_alltail<BN, Epi>(p, epi, grid, stream)));
      prof_end(stream);
      ACE_CUDA_CHECK(cudaGetLastError());
      return ACE_OK;
    }
  }
  ACE_CUDA_CHECK(launch_kernel(gemm_tc2_kernel<BN, STAGES, Epi>, dim3(grid), dim3(GEMM2_THREADS),
                               (size_t)L::TOTAL, stream, p.tma_a, p.tma_b, p.shp, epi));
  prof_end(stream);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

template <class Epi>
int launch_gemm_pair192(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  if constexpr (Epi::kHalfTile) {
    return launch_gemm_pair<192, 6, Epi>(p, epi, stream);
  } else {
    set_error("launch_gemm: BLOCK_N 192 needs a half-tile capable epilogue");
    return ACE_ERR_UNSUPPORTED;
  }
}

// bn selects the tile: 128 -> single-CTA 128 x 128 tiles, 256 / 192 -> CTA-pair 256 x bn tiles.
// (The scalar debug path always runs the epilogue per 128-column sub-tile, like both kernels do.)
template <class Epi>
int launch_gemm(const GemmPlan& p, const Epi& epi, cudaStream_t stream) {
  if (p.shp.M <= 0 || p.shp.N <= 0) return ACE_OK;
  if (gemm_debug_reference() || p.bn == 128) return launch_gemm_bn<128, 6, Epi>(p, epi, stream);
  if (p.bn == 256) return launch_gemm_pair<256, 6, Epi>(p, epi, stream);
  if (p.bn == 192) return launch_gemm_pair192<Epi>(p, epi, stream);
  set_error("launch_gemm: unsupported BLOCK_N %d", p.bn);
  return ACE_ERR_UNSUPPORTED;
}

#endif  // __CUDACC__

}  // namespace ace
