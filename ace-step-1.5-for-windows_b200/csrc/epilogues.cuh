// epilogues.cuh — fused GEMM epilogues (run on fp32 accumulator rows read from TMEM).
//
// Contract (see gemm.cuh): `run<BN>(acc, row, n0, M, N, stg)` is called by every thread of an
// epilogue warp (acc.load32 is warp-collective, so TMEM loads are never predicated); `row` is this
// thread's global output row (row - lane is the warp's first row) and may be >= M, in which case
// nothing may be stored.  `stg` is a 4 KB per-warp shared-memory slab.
//
// Memory access pattern: tcgen05.ld hands each THREAD one accumulator ROW, so storing straight from
// registers makes every 16-byte store of a warp hit 32 different 128-byte lines.  Measured on the
// DiT's o_proj GEMM that made the epilogue LSU-bound at 6.4 us — 30 % of the kernel
// (round-1 A/B measurement; the raw log was not kept — the kernel-level ncu rows of that build are in profiles/r1_v3_ncu_full_summary.txt).  The DiT epilogues therefore go through the warp's slab: the thread-per-row
// side writes/reads 16-byte chunks XOR-swizzled by row (bank-conflict free), and the global side
// moves the slab with 8 lanes per row (4 rows x 128 contiguous bytes per instruction).  The codec's
// EpiConv stays thread-per-row (measured faster there, see its comment).
//
// The roundings mirror the reference's bf16 execution where that is free: a Linear output is
// rounded to bf16 before the next elementwise op, RMSNorm rounds x*rsqrt(var) and then w*x_hat,
// RoPE rounds each product and the sum, etc. (transformers Qwen3RMSNorm / apply_rotary_pos_emb,
// imported by the reference at modeling_acestep_v15_turbo.py:30-39).
#pragma once

#include "common.cuh"

namespace ace {

#ifdef __CUDACC__

constexpr int EPI_STAGE_BYTES = 4096;  // per warp: 32 rows x 64 bf16

#ifdef ACE_GEMM_TIMING
__device__ long long g_tail_clk[32];  // probe builds: block 0 / thread 128, clock64 at the steps of the tail path
#define ACE_TCLK(i) do { if (blockIdx.x == 0 && threadIdx.x == 128) g_tail_clk[i] = clock64(); } while (0)
#else
#define ACE_TCLK(i) do { } while (0)
#endif

__device__ __forceinline__ void l1_prefetch(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
__device__ __forceinline__ void store_bf16x32(bf16* p, const float (&v)[32]) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 w;
    w.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    w.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    w.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    w.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    q[i] = w;
  }
}
__device__ __forceinline__ void load_bf16x32(const bf16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 w = q[i];
    unpack_bf16x2(w.x, v[8 * i + 0], v[8 * i + 1]);
    unpack_bf16x2(w.y, v[8 * i + 2], v[8 * i + 3]);
    unpack_bf16x2(w.z, v[8 * i + 4], v[8 * i + 5]);
    unpack_bf16x2(w.w, v[8 * i + 6], v[8 * i + 7]);
  }
}

// One warp's [32 rows x 64 bf16 columns] slab: 128-byte rows, 16-byte chunks XOR-swizzled by row.
struct WarpStage {
  uint8_t* base;
  __device__ __forceinline__ uint4* at(int row, int chunk) const {
    return reinterpret_cast<uint4*>(base + row * 128 + ((chunk ^ (row & 7)) << 4));
  }
  // thread-per-row side: columns [32*c32, 32*c32 + 32) of row `lane`
  __device__ __forceinline__ void put32(int lane, int c32, const float (&v)[32]) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 w;
      w.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
      w.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
      w.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
      w.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
      *at(lane, c32 * 4 + q) = w;
    }
  }
  __device__ __forceinline__ void get32(int lane, int c32, float (&v)[32]) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 w = *at(lane, c32 * 4 + q);
      unpack_bf16x2(w.x, v[8 * q + 0], v[8 * q + 1]);
      unpack_bf16x2(w.y, v[8 * q + 2], v[8 * q + 3]);
      unpack_bf16x2(w.z, v[8 * q + 4], v[8 * q + 5]);
      unpack_bf16x2(w.w, v[8 * q + 6], v[8 * q + 7]);
    }
  }
  // global side: 8 lanes per row, 4 rows per instruction.  `addr(r)` returns the address of slab
  // row r's first element, or nullptr when that row must be skipped; ncols (multiple of 8) = valid
  // columns of the slab.
  template <class AddrFn>
  __device__ __forceinline__ void load_rows(int lane, int ncols, AddrFn addr) const {
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      const bf16* p = addr(r);
      if (p != nullptr && ch * 8 < ncols) *at(r, ch) = *reinterpret_cast<const uint4*>(p + ch * 8);
    }
    __syncwarp();
  }
  // split version of load_rows: global -> registers now, registers -> slab later (lets the loads of
  // the next slab fly while the current one is being computed on)
  template <class AddrFn>
  __device__ __forceinline__ void fetch_rows(int lane, int ncols, AddrFn addr, uint4 (&regs)[8]) const {
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bf16* p = addr(i * 4 + (lane >> 3));
      regs[i] = (p != nullptr && ch * 8 < ncols) ? *reinterpret_cast<const uint4*>(p + ch * 8) : make_uint4(0, 0, 0, 0);
    }
  }
  __device__ __forceinline__ void commit_rows(int lane, const uint4 (&regs)[8]) const {
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) *at(i * 4 + (lane >> 3), ch) = regs[i];
    __syncwarp();
  }
  template <class AddrFn>
  __device__ __forceinline__ void store_rows(int lane, int ncols, AddrFn addr) const {
    __syncwarp();
    const int ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      bf16* p = addr(r);
      if (p != nullptr && ch * 8 < ncols) *reinterpret_cast<uint4*>(p + ch * 8) = *at(r, ch);
    }
    __syncwarp();
  }
};

// ---------------------------------------------------------------------------------------------
// Deferred RMSNorm / AdaLN (the DiT layer's three norms and the output norm run as NO kernel of their own).
//
// Reference (AceStepDiTLayer.forward :490-496, :513, :526-529; output norm :1488-1493):
//     hn = rmsnorm(h) * w * (1 + scale_b) + shift_b ;   y = hn @ W^T
// Split by linearity of the GEMM in its A operand:
//     y[m, n] = rstd[m] * sum_k (h[m,k] * c_b[k]) * W[n,k]  +  sum_k shift_b[k] * W[n,k]
//             = rstd[m] * (g @ W^T)[m, n] + bs_b[n],      c_b = w * (1 + scale_b),  g = h * c_b
//   * the PRODUCER of h (the residual-adding epilogue of the previous GEMM) also writes g = bf16(h * c) and this
//     row's sum of squares, as one fp32 partial per 64-column slab in a fixed slot (NormOut) — no atomics, the
//     consumer adds the D/64 partials in slot order, so replays are bit-identical;
//   * the CONSUMER GEMM runs on g and its epilogue applies rstd[m] and bs (NormIn);
//   * c_b and bs_b = shift_b @ W^T depend on the timestep only: they live in the handle's timestep cache
//     (dit.cu), one entry per distinct t, indexed per batch item through `slot[b]`.
// Fewer bf16 roundings than the reference's chain (one on g instead of three on hn), same fp32 statistics.
// ---------------------------------------------------------------------------------------------
struct NormOut {
  bf16* g;            // [M, ldg] A operand of the consuming GEMM; null with ssp set = statistics only (the consumer
                      // reads h itself: a norm without modulation has its weight folded into the consumer's W)
  long ldg;
  const bf16* cvec;   // c vector(s): cvec + slot[b] * c_stride  (c_stride 0: one constant vector)
  long c_stride;
  float* ssp;         // [M, nss] partial sums of squares, nss = D / 64, slot j = columns [64 j, 64 j + 64);
                      // null = this epilogue feeds no norm at all
  int nss;
  const int* slot;    // [batch] timestep-cache entry of each batch item (null with c_stride 0)
  int S;              // rows per batch item
  // called from the epilogue's prefetch hook (main loop still running): this row's c lines -> L1
  __device__ __forceinline__ void warm(int row, int n0, int M, int N) const {
    if (g == nullptr || row >= M) return;
    const bf16* c = cvec + (c_stride ? (long)__ldg(slot + row / S) * c_stride : 0) + n0;
    l1_prefetch(c);
    if (n0 + 64 < N) l1_prefetch(c + 64);
  }
  // slab `stg` holds bf16 rows [row0, row0 + 32) x columns [col, col + 64) of the new h (lane's own row = `lane`):
  // writes this row's partial, turns the slab into g in place and stores it.
  __device__ __forceinline__ void emit(const WarpStage& stg, int lane, int row, int row0, int col, int ncols,
                                       int M) const {
    if (g == nullptr) {  // statistics only
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 r = *stg.at(lane, q);
        float x0, x1;
        unpack_bf16x2(r.x, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
        unpack_bf16x2(r.y, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
        unpack_bf16x2(r.z, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
        unpack_bf16x2(r.w, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
      }
      if (row < M) ssp[(size_t)row * nss + (col >> 6)] = ss;
      return;
    }
    const int b = row < M ? row / S : 0;
    const uint4* cp = reinterpret_cast<const uint4*>(cvec + (c_stride ? (long)__ldg(slot + b) * c_stride : 0) + col);
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4* sl = stg.at(lane, q);
      uint4 r = *sl;
      const uint4 c = __ldg(cp + q);
      float x0, x1;
      unpack_bf16x2(r.x, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
      unpack_bf16x2(r.y, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
      unpack_bf16x2(r.z, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
      unpack_bf16x2(r.w, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
      r.x = bmul2(r.x, c.x); r.y = bmul2(r.y, c.y); r.z = bmul2(r.z, c.z); r.w = bmul2(r.w, c.w);
      *sl = r;
    }
    if (row < M) ssp[(size_t)row * nss + (col >> 6)] = ss;
    stg.store_rows(lane, ncols, [&](int r) -> bf16* {
      return row0 + r < M ? g + (size_t)(row0 + r) * ldg + col : nullptr;
    });
  }
};

struct NormIn {
  const float* ssp;   // null = plain GEMM (no deferred norm)
  int nss;
  float inv_d, eps;
  const float* bs;    // bias rows: bs + slot[b] * bs_stride + n; null = none (plain RMSNorm: no shift)
  long bs_stride;
  const int* slot;
  int S;
  // called from the epilogue's prefetch hook: this row's bias lines for 128 columns from n0 -> L1 (they are read
  // 32 columns at a time between TMEM loads; from L2 each of those reads is a ~700-cycle round trip)
  __device__ __forceinline__ void warm(int row, int n0, int M, int N) const {
    if (bs == nullptr || row >= M) return;
    const float* b = bs + (long)__ldg(slot + row / S) * bs_stride + n0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + 32 * i < N) l1_prefetch(b + 32 * i);
  }
  // rstd of this thread's row (called before the accumulator is ready)
  __device__ __forceinline__ float rstd(int row, int M) const {
    if (ssp == nullptr) return 1.0f;
    if (row >= M) return 0.0f;
    const float4* p = reinterpret_cast<const float4*>(ssp + (size_t)row * nss);
    float tot = 0.f;
    for (int j = 0; j < (nss >> 2); ++j) {  // fixed order: bit-identical on replay
      const float4 v = p[j];
      tot += v.x; tot += v.y; tot += v.z; tot += v.w;
    }
    return rsqrtf(tot * inv_d + eps);
  }
  __device__ __forceinline__ const float* bias_row(int row, int M) const {
    if (bs == nullptr) return nullptr;
    const int b = row < M ? row / S : 0;
    return bs + (long)__ldg(slot + b) * bs_stride;
  }
  // v[i] = v[i] * rstd + bias[n + i]  for 32 consecutive columns
  __device__ __forceinline__ void apply32(float (&v)[32], float rs, const float* brow, int n) const {
    if (ssp == nullptr) return;
    if (brow != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + n) + i);
        v[4 * i + 0] = fmaf(v[4 * i + 0], rs, b4.x); v[4 * i + 1] = fmaf(v[4 * i + 1], rs, b4.y);
        v[4 * i + 2] = fmaf(v[4 * i + 2], rs, b4.z); v[4 * i + 3] = fmaf(v[4 * i + 3], rs, b4.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= rs;
    }
  }
};

// out[row, n] = bf16(acc + bias[n])                       (proj_in, condition_embedder)
struct EpiBias {
  static constexpr bool kHalfTile = true;  // run<64> on a 64-column half tile is valid
  static constexpr bool kTmaTail = false;
  bf16* out;
  long ldo;
  const bf16* bias;  // may be null
  NormOut no = {};   // no.g != null: also feed the next norm (proj_in -> layer 0's self-attention norm)
  __device__ __forceinline__ float prefetch(int row, int n0, int M, int N, const WarpStage&) const {
    no.warm(row, n0, M, N);
    return 0.f;
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N, const WarpStage& stg,
                                      float) const {
    const int lane = threadIdx.x & 31, row0 = row - lane;
#pragma unroll 1
    for (int s = 0; s < BN; s += 64) {
      if (n0 + s >= N) break;  // warp-uniform
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        float v[32];
        acc.load32(s + c32 * 32, v);
        if (bias && n0 + s + c32 * 32 < N) {
          float b[32];
          load_bf16x32(bias + n0 + s + c32 * 32, b);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += b[i];
        }
        stg.put32(lane, c32, v);
      }
      const int ncols = N - (n0 + s) < 64 ? N - (n0 + s) : 64;
      stg.store_rows(lane, ncols, [&](int r) -> bf16* {
        return row0 + r < M ? out + (size_t)(row0 + r) * ldo + n0 + s : nullptr;
      });
      if (no.ssp != nullptr) no.emit(stg, lane, row, row0, n0 + s, ncols, M);
    }
  }
};

// out[row * ldo + n] = acc (fp32, thread-per-row: rows are few)      — the bias rows bs = shift @ W^T of the
// timestep cache (dit.cu); row r of the GEMM is cache entry first + r.
struct EpiStoreF32 {
  static constexpr bool kHalfTile = true;
  static constexpr bool kTmaTail = false;
  float* out;
  long ldo;
  __device__ __forceinline__ float prefetch(int, int, int, int, const WarpStage&) const { return 0.f; }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N, const WarpStage&, float) const {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      acc.load32(c, v);
      if (row < M && n0 + c < N) {
        float4* o = reinterpret_cast<float4*>(out + (size_t)row * ldo + n0 + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
};

// Fused q/k/v projection epilogue (BN must be 128 = head_dim so one tile is one head):
//   columns [0, nq)        : query heads  -> RMSNorm(head_dim) [+ RoPE]
//   columns [nq, nq + nk)  : key heads    -> RMSNorm(head_dim) [+ RoPE]
//   columns [nq + nk, N)   : value heads  -> plain bf16
// Follows AceStepAttention.forward (modeling_acestep_v15_turbo.py:301, 317-318, 335-340).
struct EpiQKV {
  static constexpr bool kHalfTile = false;
  static constexpr bool kTmaTail = false;
  bf16* out;
  long ldo;
  int nq, nk;
  const bf16* q_norm_w;  // [128]
  const bf16* k_norm_w;  // [128]
  const bf16* cos_tab;   // [S, 64] bf16 (cos(pos * inv_freq_i)); null => no RoPE
  const bf16* sin_tab;   // [S, 64]
  int S;                 // tokens per batch item (position = row % S)
  float eps;
  NormIn ni = {};        // ni.ssp != null: A was g = h * c, apply rstd[row] and the shift bias here (deferred AdaLN)
  // norm weights and this row's RoPE table lines -> L1 while the main loop is still running; returns rstd[row]
  __device__ __forceinline__ float prefetch(int row, int n0, int M, int N, const WarpStage&) const {
    ni.warm(row, n0, M, N);
    const float rs = ni.rstd(row, M);
    if (n0 >= nq + nk || row >= M) return rs;
    const bf16* w = (n0 < nq) ? q_norm_w : k_norm_w;
    l1_prefetch(w);
    l1_prefetch(w + 64);
    if (cos_tab != nullptr) {
      l1_prefetch(cos_tab + (size_t)(row % S) * 64);
      l1_prefetch(sin_tab + (size_t)(row % S) * 64);
    }
    return rs;
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N, const WarpStage& stg,
                                      float rs) const {
    static_assert(BN == 128, "EpiQKV needs one head per tile");
    const int lane = threadIdx.x & 31, row0 = row - lane;
    const float* brow = ni.bias_row(row, M);
    auto out_addr = [&](int col0) {
      return [=](int r) -> bf16* { return row0 + r < M ? out + (size_t)(row0 + r) * ldo + n0 + col0 : nullptr; };
    };
    if (n0 >= nq + nk) {  // value head: tile-uniform branch
#pragma unroll 1
      for (int s = 0; s < BN; s += 64) {
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
          float v[32];
          acc.load32(s + c32 * 32, v);
          ni.apply32(v, rs, brow, n0 + s + c32 * 32);
          stg.put32(lane, c32, v);
        }
        stg.store_rows(lane, 64, out_addr(s));
      }
      return;
    }
    const bf16* w = (n0 < nq) ? q_norm_w : k_norm_w;
    // pass 1: round the projection to bf16 (kept packed, 64 registers) and take its mean of squares
    uint32_t xp[64];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[32];
      acc.load32(c * 32, v);
      ni.apply32(v, rs, brow, n0 + c * 32);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t t = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        float x0, x1;
        unpack_bf16x2(t, x0, x1);
        ss = fmaf(x0, x0, ss);
        ss = fmaf(x1, x1, ss);
        xp[c * 16 + i] = t;
      }
    }
    const float rstd = rsqrtf(ss * (1.0f / 128.0f) + eps);
    const int pos = row < M ? (row % S) : 0;
    // pass 2: normalise, rotate pairs (i, i + 64) in packed bf16; word j of xp holds columns 2j, 2j+1.
    // The first 64-column half is staged and flushed while the second half waits in xp[32..63].
    const uint4* wv = reinterpret_cast<const uint4*>(w);
    const uint4* cv = reinterpret_cast<const uint4*>(cos_tab + (size_t)pos * 64);
    const uint4* sv = reinterpret_cast<const uint4*>(sin_tab + (size_t)pos * 64);
    auto norm2 = [&](uint32_t x, uint32_t wgt) {  // bf16(w * bf16(x * rstd)) on two lanes
      float x0, x1;
      unpack_bf16x2(x, x0, x1);
      return bmul2(wgt, pack_bf16x2(x0 * rstd, x1 * rstd));
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) {  // 8 columns of each half per iteration
      const uint4 wl = __ldg(wv + q), wh = __ldg(wv + 8 + q);
      const uint32_t wls[4] = {wl.x, wl.y, wl.z, wl.w}, whs[4] = {wh.x, wh.y, wh.z, wh.w};
      uint32_t lo[4], hi[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        lo[k] = norm2(xp[4 * q + k], wls[k]);
        hi[k] = norm2(xp[32 + 4 * q + k], whs[k]);
      }
      if (cos_tab != nullptr) {
        const uint4 c4 = __ldg(cv + q), s4 = __ldg(sv + q);
        const uint32_t cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // q*cos + rotate_half(q)*sin, every product and the sum rounded to bf16
          const uint32_t a = bsub2(bmul2(lo[k], cs[k]), bmul2(hi[k], sn[k]));
          const uint32_t b = badd2(bmul2(hi[k], cs[k]), bmul2(lo[k], sn[k]));
          lo[k] = a;
          hi[k] = b;
        }
      }
      *stg.at(lane, q) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) xp[32 + 4 * q + k] = hi[k];
    }
    stg.store_rows(lane, 64, out_addr(0));
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *stg.at(lane, q) = make_uint4(xp[32 + 4 * q], xp[33 + 4 * q], xp[34 + 4 * q], xp[35 + 4 * q]);
    stg.store_rows(lane, 64, out_addr(64));
  }
};

// h[row, n] = bf16(h + bf16(bf16(acc) * gate[b, n]))   (gate == null: plain residual)
// AceStepDiTLayer.forward lines 508, 523, 530.
struct EpiGatedResid {
  static constexpr bool kHalfTile = true;
  // The CTA-pair kernel runs the LAST tile of every CTA (every tile of a single-wave problem: M = 1500 at N = 2048)
  // through tail_box() instead of run(): the residual tile arrives by TMA in the operand ring's freed slots while
  // the main loop drains, the update happens in place in shared memory, and h (and g) leave by TMA stores — the
  // epilogue warps only read TMEM and touch shared memory.  Multi-wave problems (M = 3000 / 6000) run EVERY tile
  // that way in the ALLTAIL variant of the kernel (dedicated residual boxes, tail_box<false> + tail_g: gemm.cuh).
  // Needs the tensor maps below (use_tma); run() is the path without them (condition encoder, probe A/B).
  static constexpr bool kTmaTail = true;
  bf16* h;  // read-modify-write in place
  long ldh;
  const bf16* gate;  // gate vector of batch item b at gate + gsel(b) * gate_ld, gsel(b) = slot ? slot[b] : b; or null
  long gate_ld;
  int S;  // rows per batch item
  const int* slot = nullptr;  // [batch] timestep-cache entry per batch item (the DiT); null: gate is [batch, gate_ld]
  NormOut no = {};            // no.g != null: also feed the next norm
  int use_tma = 0;            // tm_h / tm_g valid: [N cols, M rows] bf16 maps over h and no.g, box [64 x 128], SWIZZLE_128B
  CUtensorMap tm_h = {};
  CUtensorMap tm_g = {};
  // One 64-column box of the tail path.  `hbox`: this CTA's [128 rows x 64 cols] residual box as TMA wrote it
  // (128-byte rows, 16-byte chunks XOR-swizzled by row & 7), updated IN PLACE to the new h; `gbox`: same layout,
  // receives g.  `r` = this thread's row inside the CTA tile, `row` its global row, `col` the box's first column.
  // Called by every tail-path thread while the main loop is still running: resolves this row's gate / c vectors
  // (the timestep-cache indirection is a dependent L2 round trip — about 900 cycles that must not sit between the
  // accumulator and the stores) and pulls their lines for columns [col0, col0 + ncols) into L1.
  struct TailPre {
    const bf16* grow;  // this row's gate vector (column 0) or null
    const bf16* crow;  // this row's c vector (column 0) or null
  };
  __device__ __forceinline__ TailPre tail_prefetch(int row, int M, int N, int col0, int ncols) const {
    TailPre t{nullptr, nullptr};
    if (gate != nullptr) t.grow = gate + (size_t)gsel(row, M) * gate_ld;
    if (no.g != nullptr) {
      const int b = row < M ? row / S : 0;
      t.crow = no.cvec + (no.c_stride ? (long)__ldg(no.slot + b) * no.c_stride : 0);
    }
    for (int c = col0; c < col0 + ncols && c < N; c += 64) {
      if (t.grow != nullptr) l1_prefetch(t.grow + c);
      if (t.crow != nullptr) l1_prefetch(t.crow + c);
    }
    return t;
  }
  // WITH_G = false: h (and the sum of squares) only; g is then made in place by tail_g() once h's store has read
  // the box (the every-tile variant of the pair kernel has no second box to put it in).
  template <bool WITH_G = true, class Acc>
  __device__ __forceinline__ void tail_box(const Acc& acc, int acc_col, int r, int row, int col, int M,
                                           uint8_t* hbox, uint8_t* gbox, const TailPre& pre) const {
    uint32_t raw[2][32];
    ACE_TCLK(16);
    acc.load32_nowait(acc_col, raw[0]);
    acc.load32_nowait(acc_col + 32, raw[1]);
    const uint4* gp = reinterpret_cast<const uint4*>(pre.grow + col);
    const uint4* cp = reinterpret_cast<const uint4*>(pre.crow + col);
    uint4 gv[8], cv[8];
    if (gate != nullptr) {
#pragma unroll
      for (int q = 0; q < 8; ++q) gv[q] = __ldg(gp + q);
    }
    if (WITH_G && no.g != nullptr) {
#pragma unroll
      for (int q = 0; q < 8; ++q) cv[q] = __ldg(cp + q);
    }
    ACE_TCLK(17);
    tmem_ld_wait();
    ACE_TCLK(18);
    float ss = 0.f;
    const uint32_t rowoff = (uint32_t)r * 128u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t off = rowoff + (uint32_t)((q ^ (r & 7)) << 4);
      uint4* hs = reinterpret_cast<uint4*>(hbox + off);
      uint4 res = *hs;
      const uint32_t* a = &raw[q >> 2][(q & 3) * 8];
      uint4 pv;
      pv.x = pack_bf16x2(__uint_as_float(a[0]), __uint_as_float(a[1]));
      pv.y = pack_bf16x2(__uint_as_float(a[2]), __uint_as_float(a[3]));
      pv.z = pack_bf16x2(__uint_as_float(a[4]), __uint_as_float(a[5]));
      pv.w = pack_bf16x2(__uint_as_float(a[6]), __uint_as_float(a[7]));
      if (gate != nullptr) {
        pv.x = bmul2(pv.x, gv[q].x); pv.y = bmul2(pv.y, gv[q].y);
        pv.z = bmul2(pv.z, gv[q].z); pv.w = bmul2(pv.w, gv[q].w);
      }
      res.x = badd2(res.x, pv.x); res.y = badd2(res.y, pv.y);
      res.z = badd2(res.z, pv.z); res.w = badd2(res.w, pv.w);
      *hs = res;
      if (no.ssp != nullptr) {
        float x0, x1;
        unpack_bf16x2(res.x, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
        unpack_bf16x2(res.y, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
        unpack_bf16x2(res.z, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
        unpack_bf16x2(res.w, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
      }
      if (WITH_G && no.g != nullptr) {
        uint4 gq;
        gq.x = bmul2(res.x, cv[q].x); gq.y = bmul2(res.y, cv[q].y);
        gq.z = bmul2(res.z, cv[q].z); gq.w = bmul2(res.w, cv[q].w);
        *reinterpret_cast<uint4*>(gbox + off) = gq;
      }
    }
    ACE_TCLK(19);
    if (no.ssp != nullptr && row < M) no.ssp[(size_t)row * no.nss + (col >> 6)] = ss;
  }
  // g = h * c in place over a box that holds the NEW h (same products as tail_box / run make)
  __device__ __forceinline__ void tail_g(int r, int col, uint8_t* hbox, const TailPre& pre) const {
    const uint4* cp = reinterpret_cast<const uint4*>(pre.crow + col);
    const uint32_t rowoff = (uint32_t)r * 128u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4* hs = reinterpret_cast<uint4*>(hbox + rowoff + (uint32_t)((q ^ (r & 7)) << 4));
      const uint4 v = *hs, c = __ldg(cp + q);
      *hs = make_uint4(bmul2(v.x, c.x), bmul2(v.y, c.y), bmul2(v.z, c.z), bmul2(v.w, c.w));
    }
  }
  __device__ __forceinline__ int gsel(int row, int M) const {
    const int b = row < M ? row / S : 0;
    return slot != nullptr ? __ldg(slot + b) : b;
  }
  // called BEFORE the accumulator is ready (the main loop is still running): stage the first
  // residual slab so its L2 latency is off the critical path
  __device__ __forceinline__ float prefetch(int row, int n0, int M, int N, const WarpStage& stg) const {
    const int lane = threadIdx.x & 31, row0 = row - lane;
    if (n0 >= N) return 0.f;
    const int ncols = N - n0 < 64 ? N - n0 : 64;
    if (gate != nullptr && row < M) {  // this row's gate vector: pulled into L1 for the loads in run()
      const bf16* g = gate + (size_t)gsel(row, M) * gate_ld + n0;
      l1_prefetch(g);
      if (n0 + 64 < N) l1_prefetch(g + 64);
    }
    no.warm(row, n0, M, N);
    stg.load_rows(lane, ncols, [&](int r) -> const bf16* {
      return row0 + r < M ? h + (size_t)(row0 + r) * ldh + n0 : nullptr;
    });
    return 0.f;
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N, const WarpStage& stg,
                                      float) const {
    const int lane = threadIdx.x & 31, row0 = row - lane;
    const int b = gate != nullptr ? gsel(row, M) : 0;
    const bf16* crow = nullptr;  // this row's c vector (the next norm's g = h * c), resolved once
    if (no.g != nullptr)
      crow = no.cvec + (no.c_stride ? (long)__ldg(no.slot + (row < M ? row / S : 0)) * no.c_stride : 0);
    uint4 nxt[8];
#pragma unroll 1
    for (int s = 0; s < BN; s += 64) {
      if (n0 + s >= N) break;  // warp-uniform
      const int ncols = N - (n0 + s) < 64 ? N - (n0 + s) : 64;
      auto addr = [&](int r) -> bf16* { return row0 + r < M ? h + (size_t)(row0 + r) * ldh + n0 + s : nullptr; };
      // NormOut: the sum of squares comes straight from the registers that hold the new h; g = h * c is made in
      // place in the slab once h's copy has left it (holding g in registers instead spilled: 168-register budget)
      float ss = 0.f;
      const bool has_next = s + 64 < BN && n0 + s + 64 < N;
      if (has_next) {  // next residual slab: global -> registers while this slab is processed
        const int nn = N - (n0 + s + 64) < 64 ? N - (n0 + s + 64) : 64;
        stg.fetch_rows(lane, nn, [&](int r) -> const bf16* {
          return row0 + r < M ? h + (size_t)(row0 + r) * ldh + n0 + s + 64 : nullptr;
        }, nxt);
      }
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        float v[32];
        acc.load32(s + c32 * 32, v);
        const bool gl = gate != nullptr && row < M && n0 + s + c32 * 32 < N;
        const uint4* gp = reinterpret_cast<const uint4*>(gate + (size_t)b * gate_ld + n0 + s + c32 * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4* slot = stg.at(lane, c32 * 4 + q);
          uint4 r = *slot;
          uint4 pv;
          pv.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
          pv.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
          pv.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
          pv.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
          if (gate != nullptr) {
            const uint4 g = gl ? __ldg(gp + q) : make_uint4(0, 0, 0, 0);
            pv.x = bmul2(pv.x, g.x); pv.y = bmul2(pv.y, g.y);
            pv.z = bmul2(pv.z, g.z); pv.w = bmul2(pv.w, g.w);
          }
          r.x = badd2(r.x, pv.x); r.y = badd2(r.y, pv.y);
          r.z = badd2(r.z, pv.z); r.w = badd2(r.w, pv.w);
          *slot = r;
          if (no.ssp != nullptr) {
            float x0, x1;
            unpack_bf16x2(r.x, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
            unpack_bf16x2(r.y, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
            unpack_bf16x2(r.z, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
            unpack_bf16x2(r.w, x0, x1); ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
          }
        }
      }
      stg.store_rows(lane, ncols, addr);
      if (no.ssp != nullptr && row < M) no.ssp[(size_t)row * no.nss + ((n0 + s) >> 6)] = ss;
      if (no.g != nullptr) {
        const uint4* cp = reinterpret_cast<const uint4*>(crow + n0 + s);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4* sl = stg.at(lane, q);
          const uint4 r = *sl, c = __ldg(cp + q);
          *sl = make_uint4(bmul2(r.x, c.x), bmul2(r.y, c.y), bmul2(r.z, c.z), bmul2(r.w, c.w));
        }
        stg.store_rows(lane, ncols, [&](int r) -> bf16* {
          return row0 + r < M ? no.g + (size_t)(row0 + r) * no.ldg + n0 + s : nullptr;
        });
      }
      if (has_next) stg.commit_rows(lane, nxt);
    }
  }
};

// SwiGLU: B is packed so that tile columns [0,64) are gate features f0..f0+63 and [64,128) the
// matching up features; out[row, f0 + i] = bf16(bf16(silu(g)) * u).   (Qwen3MLP.forward)
struct EpiSwiGLU {
  static constexpr bool kHalfTile = false;
  static constexpr bool kTmaTail = false;
  bf16* out;
  long ldo;
  NormIn ni = {};  // deferred AdaLN of the MLP norm (see NormIn)
  __device__ __forceinline__ float prefetch(int row, int n0, int M, int N, const WarpStage&) const {
    ni.warm(row, n0, M, N);
    return ni.rstd(row, M);
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N, const WarpStage& stg,
                                      float rs) const {
    static_assert(BN == 128, "EpiSwiGLU packs 64 gate + 64 up columns per tile");
    const int lane = threadIdx.x & 31, row0 = row - lane;
    const int f0 = (n0 >> 7) << 6;
    const float* brow = ni.bias_row(row, M);
#pragma unroll
    for (int c32 = 0; c32 < 2; ++c32) {
      float g[32], u[32];
      acc.load32(c32 * 32, g);
      acc.load32(c32 * 32 + 64, u);
      ni.apply32(g, rs, brow, n0 + c32 * 32);
      ni.apply32(u, rs, brow, n0 + c32 * 32 + 64);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float x0, x1;
          unpack_bf16x2(pack_bf16x2(g[8 * q + 2 * k], g[8 * q + 2 * k + 1]), x0, x1);
          const uint32_t sl = pack_bf16x2(__fdividef(x0, 1.0f + __expf(-x0)), __fdividef(x1, 1.0f + __expf(-x1)));
          o[k] = bmul2(sl, pack_bf16x2(u[8 * q + 2 * k], u[8 * q + 2 * k + 1]));
        }
        *stg.at(lane, c32 * 4 + q) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    stg.store_rows(lane, 64, [&](int r) -> bf16* {
      return row0 + r < M ? out + (size_t)(row0 + r) * ldo + f0 : nullptr;
    });
  }
};

// proj_out (ConvTranspose1d k=2 s=2 as a GEMM with N = 2*64): column n = k*64 + o lands at
// vt[b, 2*s + k, o]; frames >= T (the odd-length pad) are cropped.  (turbo modeling :1284-1294,1498)
struct EpiProjOut {
  static constexpr bool kHalfTile = false;
  static constexpr bool kTmaTail = false;
  bf16* vt;  // [Bc, T, 64]
  const bf16* bias;  // [64]
  int S, T;
  NormIn ni = {};  // deferred output AdaLN (:1488-1493)
  __device__ __forceinline__ float prefetch(int row, int n0, int M, int N, const WarpStage&) const {
    ni.warm(row, n0, M, N);
    return ni.rstd(row, M);
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N, const WarpStage& stg,
                                      float rs) const {
    static_assert(BN == 128, "EpiProjOut");
    const int lane = threadIdx.x & 31, row0 = row - lane;
    const float* brow = ni.bias_row(row, M);
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {  // kernel tap k = 64-column half k
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        float v[32], bb[32];
        acc.load32(k * 64 + c32 * 32, v);
        ni.apply32(v, rs, brow, n0 + k * 64 + c32 * 32);
        load_bf16x32(bias + c32 * 32, bb);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += bb[i];
        stg.put32(lane, c32, v);
      }
      stg.store_rows(lane, 64, [&](int r) -> bf16* {
        const int gr = row0 + r;
        if (gr >= M) return nullptr;
        const int b = gr / S, t = 2 * (gr % S) + k;
        return t < T ? vt + ((size_t)b * T + t) * 64 : nullptr;
      });
    }
  }
};

// snake(x) = x + 1/(exp(beta)+1e-9) * sin^2(exp(alpha) * x); a = exp(alpha), ib = 1/(exp(beta)+1e-9)
// precomputed per channel at pack time.  sin^2(y) = 0.5 - 0.5 cos(2y), argument reduced to [-pi, pi]
// before the MUFU cosine.  (acestep/models/mlx/vae_model.py:24-56)
__device__ __forceinline__ float snake_f(float x, float a, float ib) {
  float y = 2.0f * a * x;
  // round(y / 2pi) by the add-and-subtract-1.5*2^23 trick (exact round-to-nearest-even for |y / 2pi| < 2^22;
  // larger arguments lose all phase information in fp32 anyway): FRND would be a second XU-pipe
  // instruction per element next to the cosine, and the XU pipe is what paces the codec epilogues
  float k = y * 0.15915494309189535f;
  k = (k + 12582912.0f) - 12582912.0f;
  y = fmaf(-6.283185307179586f, k, y);
  return fmaf(ib, 0.5f - 0.5f * __cosf(y), x);
}

// Codec convolution epilogue.  v = bf16(acc + bias[c]) (+ resid) -> optional main store,
// optional Snake'd copy for the next conv's A operand.
// Flat element index = row*ldo + n + off must lie in [0, total) — this is how the transposed
// convolution's "-padding" shift and its ragged ends are cropped.
// Thread-per-row on purpose: the codec's rows are short (128..2048 channels) and its GEMMs have few
// k-blocks, so the epilogue's latency is exposed; the staged/coalesced variant measured 11 % slower
// on the whole decode (round-1 A/B measurement; the raw log was not kept — the kernel-level ncu rows of that build are in profiles/r1_v3_ncu_full_summary.txt), the 16-byte-per-thread stores merge in L2.
struct EpiConv {
  static constexpr bool kHalfTile = true;  // run<64> on a 64-column half tile is valid
  static constexpr bool kTmaTail = false;
  bf16* out_main;        // may be null
  bf16* out_snake;       // may be null
  const bf16* resid;     // may be null; same indexing as out
  const float* bias;     // [chan_mod] or null
  const float* sn_a;     // [chan_mod] (used when out_snake)
  const float* sn_ib;    // [chan_mod]
  long ldo, off, total;
  int chan_mod;          // channel of column n is n % chan_mod (a multiple of 32)
  // residual rows are re-read by the thread that produces the output row: pull them into L1 early
  __device__ __forceinline__ float prefetch(int row, int n0, int M, int N, const WarpStage&) const {
    if (resid == nullptr || row >= M) return 0.f;
    const long idx = (long)row * ldo + n0 + off;
    if (idx < 0 || idx >= total) return 0.f;
    l1_prefetch(resid + idx);
    l1_prefetch(resid + idx + 64);
    return 0.f;
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N, const WarpStage&, float) const {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      const long idx = (long)row * ldo + n0 + c + off;
      const bool ok = row < M && n0 + c < N && idx >= 0 && idx < total;
      const int ch = (n0 + c) % chan_mod;
      // issue the global loads first so they overlap the TMEM read
      uint4 rq[4];
      if (ok && resid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rq[i] = reinterpret_cast<const uint4*>(resid + idx)[i];
      }
      float v[32];
      acc.load32(c, v);
      if (ok) {
        if (bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + ch) + i);
            v[4 * i + 0] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
          }
        }
        // every rounding two elements at a time (F2FP.PACK_AB): the bf16 conversions share the XU pipe
        // with the Snake cosine and paced the memory-bound codec stages; badd2 is an exactly rounded bf16 add
        uint32_t xp[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) xp[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        if (resid) {
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(rq);
#pragma unroll
          for (int i = 0; i < 16; ++i) xp[i] = badd2(xp[i], rw[i]);
        }
        if (out_main) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<uint4*>(out_main + idx)[i] = make_uint4(xp[4 * i], xp[4 * i + 1], xp[4 * i + 2], xp[4 * i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) unpack_bf16x2(xp[i], v[2 * i], v[2 * i + 1]);
        if (out_snake) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(sn_a + ch) + i);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(sn_ib + ch) + i);
            v[4 * i + 0] = snake_f(v[4 * i + 0], a4.x, b4.x);
            v[4 * i + 1] = snake_f(v[4 * i + 1], a4.y, b4.y);
            v[4 * i + 2] = snake_f(v[4 * i + 2], a4.z, b4.z);
            v[4 * i + 3] = snake_f(v[4 * i + 3], a4.w, b4.w);
          }
          store_bf16x32(out_snake + idx, v);
        }
      }
    }
  }
};

#endif  // __CUDACC__

}  // namespace ace
