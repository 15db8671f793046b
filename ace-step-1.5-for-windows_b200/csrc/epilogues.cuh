// epilogues.cuh — fused GEMM epilogues (run on fp32 accumulator rows read from TMEM).
//
// Contract (see gemm.cuh): `run<BN>(acc, row, n0, M, N)` is called by every thread of an epilogue
// warp (acc.load32 is warp-collective, so loads are never predicated); `row` is this thread's
// global output row and may be >= M, in which case nothing may be stored.
//
// The roundings mirror the reference's bf16 execution where that is free: a Linear output is
// rounded to bf16 before the next elementwise op, RMSNorm rounds x*rsqrt(var) and then w*x_hat,
// RoPE rounds each product and the sum, etc. (transformers Qwen3RMSNorm / apply_rotary_pos_emb,
// imported by the reference at modeling_acestep_v15_turbo.py:30-39).
#pragma once

#include "common.cuh"

namespace ace {

#ifdef __CUDACC__

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

// out[row, n] = bf16(acc + bias[n])                       (proj_in, condition_embedder)
struct EpiBias {
  static constexpr bool kHalfTile = true;  // run<64> on a 64-column half tile is valid
  bf16* out;
  long ldo;
  const bf16* bias;  // may be null
  __device__ __forceinline__ void prefetch(int, int, int, int) const {}
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N) const {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      acc.load32(c, v);
      if (row < M && n0 + c < N) {
        if (bias) {
          float b[32];
          load_bf16x32(bias + n0 + c, b);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += b[i];
        }
        store_bf16x32(out + (size_t)row * ldo + n0 + c, v);
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
  static constexpr bool kHalfTile = false;  // run<64> on a 64-column half tile is valid
  bf16* out;
  long ldo;
  int nq, nk;
  const bf16* q_norm_w;  // [128]
  const bf16* k_norm_w;  // [128]
  const bf16* cos_tab;   // [S, 64] bf16 (cos(pos * inv_freq_i)); null => no RoPE
  const bf16* sin_tab;   // [S, 64]
  int S;                 // tokens per batch item (position = row % S)
  float eps;
  __device__ __forceinline__ void prefetch(int, int, int, int) const {}
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N) const {
    static_assert(BN == 128, "EpiQKV needs one head per tile");
    const bool ok = row < M;
    if (n0 >= nq + nk) {  // value head: tile-uniform branch
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        acc.load32(c, v);
        if (ok) store_bf16x32(out + (size_t)row * ldo + n0 + c, v);
      }
      return;
    }
    const bf16* w = (n0 < nq) ? q_norm_w : k_norm_w;
    // pass 1: mean of squares of the bf16-rounded projection
    float ss = 0.f;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      acc.load32(c, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = bf16_round(v[i]);
        ss += x * x;
      }
    }
    const float rstd = rsqrtf(ss * (1.0f / 128.0f) + eps);
    const int pos = ok ? (row % S) : 0;
    // pass 2: normalise, rotate pairs (i, i + 64)
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      float lo[32], hi[32], wl[32], wh[32];
      acc.load32(c, lo);
      acc.load32(c + 64, hi);
      load_bf16x32(w + c, wl);
      load_bf16x32(w + c + 64, wh);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        lo[i] = bf16_round(wl[i] * bf16_round(bf16_round(lo[i]) * rstd));
        hi[i] = bf16_round(wh[i] * bf16_round(bf16_round(hi[i]) * rstd));
      }
      if (cos_tab != nullptr) {
        float cs[32], sn[32];
        load_bf16x32(cos_tab + (size_t)pos * 64 + c, cs);
        load_bf16x32(sin_tab + (size_t)pos * 64 + c, sn);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          // q*cos + rotate_half(q)*sin, every product and the sum rounded to bf16
          float a = bf16_round(bf16_round(lo[i] * cs[i]) + bf16_round(-hi[i] * sn[i]));
          float b = bf16_round(bf16_round(hi[i] * cs[i]) + bf16_round(lo[i] * sn[i]));
          lo[i] = a;
          hi[i] = b;
        }
      }
      if (ok) {
        store_bf16x32(out + (size_t)row * ldo + n0 + c, lo);
        store_bf16x32(out + (size_t)row * ldo + n0 + c + 64, hi);
      }
    }
  }
};

// h[row, n] = bf16(h + bf16(bf16(acc) * gate[b, n]))   (gate == null: plain residual)
// AceStepDiTLayer.forward lines 508, 523, 530.
struct EpiGatedResid {
  static constexpr bool kHalfTile = true;  // run<64> on a 64-column half tile is valid
  bf16* h;  // read-modify-write in place
  long ldh;
  const bf16* gate;  // [Bc, gate_ld] or null
  long gate_ld;
  int S;  // rows per batch item
  // pull this thread's residual row segment (and its gate vector) into L1 ahead of the accumulator
  __device__ __forceinline__ void prefetch(int row, int n0, int M, int N) const {
    if (row >= M) return;
    const bf16* hp = h + (size_t)row * ldh + n0;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(hp));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(hp + 64));
    if (gate) {
      const bf16* gp = gate + (size_t)(row / S) * gate_ld + n0;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(gp));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(gp + 64));
    }
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N) const {
    const bool ok = row < M;
    const int b = ok ? row / S : 0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      acc.load32(c, v);
      if (ok && n0 + c < N) {
        float r[32];
        bf16* hp = h + (size_t)row * ldh + n0 + c;
        load_bf16x32(hp, r);
        if (gate) {
          float g[32];
          load_bf16x32(gate + (size_t)b * gate_ld + n0 + c, g);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = r[i] + bf16_round(bf16_round(v[i]) * g[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = r[i] + bf16_round(v[i]);
        }
        store_bf16x32(hp, v);
      }
    }
  }
};

// SwiGLU: B is packed so that tile columns [0,64) are gate features f0..f0+63 and [64,128) the
// matching up features; out[row, f0 + i] = bf16(bf16(silu(g)) * u).   (Qwen3MLP.forward)
struct EpiSwiGLU {
  static constexpr bool kHalfTile = false;  // run<64> on a 64-column half tile is valid
  bf16* out;
  long ldo;
  __device__ __forceinline__ void prefetch(int, int, int, int) const {}
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N) const {
    static_assert(BN == 128, "EpiSwiGLU packs 64 gate + 64 up columns per tile");
    const int f0 = (n0 >> 7) << 6;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      float g[32], u[32];
      acc.load32(c, g);
      acc.load32(c + 64, u);
      if (row < M) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = bf16_round(g[i]);
          float s = bf16_round(x / (1.0f + __expf(-x)));
          g[i] = s * bf16_round(u[i]);
        }
        store_bf16x32(out + (size_t)row * ldo + f0 + c, g);
      }
    }
  }
};

// proj_out (ConvTranspose1d k=2 s=2 as a GEMM with N = 2*64): column n = k*64 + o lands at
// vt[b, 2*s + k, o]; frames >= T (the odd-length pad) are cropped.  (turbo modeling :1284-1294,1498)
struct EpiProjOut {
  static constexpr bool kHalfTile = false;  // run<64> on a 64-column half tile is valid
  bf16* vt;  // [Bc, T, 64]
  const bf16* bias;  // [64]
  int S, T;
  __device__ __forceinline__ void prefetch(int, int, int, int) const {}
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N) const {
    static_assert(BN == 128, "EpiProjOut");
    const bool ok = row < M;
    const int b = ok ? row / S : 0;
    const int s = ok ? row % S : 0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32], bb[32];
      acc.load32(c, v);
      const int k = c >> 6;
      const int t = 2 * s + k;
      if (ok && t < T) {
        load_bf16x32(bias + (c & 63), bb);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += bb[i];
        store_bf16x32(vt + ((size_t)b * T + t) * 64 + (c & 63), v);
      }
    }
  }
};

// snake(x) = x + 1/(exp(beta)+1e-9) * sin^2(exp(alpha) * x); a = exp(alpha), ib = 1/(exp(beta)+1e-9)
// precomputed per channel at pack time.  sin^2(y) = 0.5 - 0.5 cos(2y), argument reduced to [-pi, pi]
// before the MUFU cosine.  (acestep/models/mlx/vae_model.py:24-56)
__device__ __forceinline__ float snake_f(float x, float a, float ib) {
  float y = 2.0f * a * x;
  y = fmaf(-6.283185307179586f, rintf(y * 0.15915494309189535f), y);
  return fmaf(ib, 0.5f - 0.5f * __cosf(y), x);
}

// Codec convolution epilogue.  v = bf16(acc + bias[c]) (+ resid) -> optional main store,
// optional Snake'd copy for the next conv's A operand, optional fp32 store.
// Flat element index = row*ldo + n + off must lie in [0, total) — this is how the transposed
// convolution's "-padding" shift and its ragged ends are cropped.
struct EpiConv {
  static constexpr bool kHalfTile = true;  // run<64> on a 64-column half tile is valid
  bf16* out_main;        // may be null
  bf16* out_snake;       // may be null
  const bf16* resid;     // may be null; same indexing as out
  const float* bias;     // [chan_mod] or null
  const float* sn_a;     // [chan_mod] (used when out_snake)
  const float* sn_ib;    // [chan_mod]
  float* out_f32;        // may be null
  long ldo, off, total;
  int chan_mod;          // channel of column n is n % chan_mod (a multiple of 32)
  // residual rows are re-read by the thread that produces the output row: pull them into L1 early
  __device__ __forceinline__ void prefetch(int row, int n0, int M, int N) const {
    if (resid == nullptr || row >= M) return;
    const long idx = (long)row * ldo + n0 + off;
    if (idx < 0 || idx >= total) return;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(resid + idx));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(resid + idx + 64));
  }
  template <int BN, class Acc>
  __device__ __forceinline__ void run(const Acc& acc, int row, int n0, int M, int N) const {
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
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
        if (resid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float r[8];
            unpack_bf16x2(rq[i].x, r[0], r[1]); unpack_bf16x2(rq[i].y, r[2], r[3]);
            unpack_bf16x2(rq[i].z, r[4], r[5]); unpack_bf16x2(rq[i].w, r[6], r[7]);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[8 * i + k] = bf16_round(v[8 * i + k] + r[k]);
          }
        }
        if (out_main) store_bf16x32(out_main + idx, v);
        if (out_f32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) out_f32[idx + i] = v[i];
        }
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
