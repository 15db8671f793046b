// attention.cu — fused softmax attention for the DiT (head_dim 128, GQA, bf16 in / fp32 softmax).
//
// One kernel covers the three attention shapes of AceStepDiTLayer
// (modeling_acestep_v15_turbo.py:286-368, masks :1405-1437):
//   * full bidirectional self-attention           (window < 0, Skv == Sq)
//   * +-W sliding-window self-attention            (|i - j| <= W; only the band's KV blocks are
//                                                   visited — no S x S mask tensor exists)
//   * cross-attention over the cached condition KV (window < 0, Skv == E)
// Q/K/V are read in place from the token-major buffers the QKV GEMM epilogue wrote
// ([tokens, heads*128], head h at column h*128), O is written token-major for the o_proj GEMM.
//
// Round-1 implementation: flash-attention style tiling (64 query rows x 64 keys per step),
// cp.async double-buffered K/V tiles in XOR-swizzled shared memory, ldmatrix + mma.sync
// m16n8k16 (legacy tensor path), online softmax with quad shuffles.  Attention is 7-13 % of the
// step FLOPs; the tcgen05/TMEM version is the planned replacement.
#include "common.cuh"
#include "kernels.h"

namespace ace {

namespace {

constexpr int HD = 128;
constexpr int BM = 64;
constexpr int BN = 64;
constexpr int ATT_THREADS = 128;
constexpr int TILE_BYTES = 64 * HD * 2;  // 16 KB

__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)(row * 256 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// loads a [64 x 128] bf16 tile (rows row0.. of a [nrows, ld] matrix) into swizzled smem
__device__ __forceinline__ void load_tile(uint32_t smem_base, const bf16* g, long ld, int row0,
                                          int nrows, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = tid + i * ATT_THREADS;  // 0..1023
    const int r = idx >> 4, c = idx & 15;
    const int gr = row0 + r;
    const bool ok = gr >= 0 && gr < nrows;
    const bf16* src = g + (long)(ok ? gr : 0) * ld + c * 8;
    cp_async16(smem_base + swz(r, c), src, ok);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + TILE_BYTES;
  const uint32_t sV = sK + 2 * TILE_BYTES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  pdl_trigger();
  pdl_wait();
  const int q0 = blockIdx.x * BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / p.group;

  const bf16* Q = p.q + (long)b * p.Sq * p.ldq + (long)h * HD;
  const bf16* K = p.k + (long)b * p.Skv * p.ldk + (long)hk * HD;
  const bf16* V = p.v + (long)b * p.Skv * p.ldv + (long)hk * HD;
  bf16* O = p.o + (long)b * p.Sq * p.ldo + (long)h * HD;

  // key range visited by this query block
  int j_lo = 0, j_hi = p.Skv;
  if (p.window >= 0) {
    j_lo = q0 - p.window;
    if (j_lo < 0) j_lo = 0;
    j_lo = (j_lo / BN) * BN;
    j_hi = q0 + BM - 1 + p.window + 1;
    if (j_hi > p.Skv) j_hi = p.Skv;
  }
  const int nblk = (j_hi - j_lo + BN - 1) / BN;

  load_tile(sQ, Q, p.ldq, q0, p.Sq, tid);
  load_tile(sK, K, p.ldk, j_lo, p.Skv, tid);
  load_tile(sV, V, p.ldv, j_lo, p.Skv, tid);
  cp_async_commit();

  uint32_t qf[8][4];
  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const int row_a = q0 + warp * 16 + g;  // this thread's two query rows: row_a, row_a + 8

  for (int blk = 0; blk < nblk; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < nblk) {
      load_tile(sK + (buf ^ 1) * TILE_BYTES, K, p.ldk, j_lo + (blk + 1) * BN, p.Skv, tid);
      load_tile(sV + (buf ^ 1) * TILE_BYTES, V, p.ldv, j_lo + (blk + 1) * BN, p.Skv, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (blk == 0) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        ldsm_x4(sQ + swz(warp * 16 + (lane & 15), 2 * ks + (lane >> 4)), qf[ks][0], qf[ks][1],
                qf[ks][2], qf[ks][3]);
    }
    const uint32_t kb = sK + buf * TILE_BYTES, vb = sV + buf * TILE_BYTES;

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int mi = lane >> 3;
        ldsm_x4(kb + swz(np * 16 + (mi >> 1) * 8 + (lane & 7), 2 * ks + (mi & 1)), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qf[ks], b0, b1);
        mma_bf16(s[2 * np + 1], qf[ks], b2, b3);
      }
    }

    // mask + online softmax (scores scaled into the log2 domain)
    const int jb = j_lo + blk * BN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = jb + nt * 8 + 2 * tq + (e & 1);
        const int i = row_a + (e >> 1) * 8;
        bool ok = j < p.Skv;
        if (p.window >= 0) {
          const int d = i - j;
          ok = ok && d <= p.window && d >= -p.window;
        }
        const float v = ok ? s[nt][e] * p.scale_log2 : -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      msafe[r] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = exp2f(m_run[r] - msafe[r]);  // m_run == -inf -> 0
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - msafe[0]);
      const float p1 = exp2f(s[nt][1] - msafe[0]);
      const float p2 = exp2f(s[nt][2] - msafe[1]);
      const float p3 = exp2f(s[nt][3] - msafe[1]);
      // probabilities go to bf16 for P @ V (as in the reference); the row sum stays fp32
      const uint32_t w01 = pack_bf16x2(p0, p1), w23 = pack_bf16x2(p2, p3);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = w01;
      pf[nt >> 1][(nt & 1) * 2 + 1] = w23;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      o[nt][0] *= corr[0];
      o[nt][1] *= corr[0];
      o[nt][2] *= corr[1];
      o[nt][3] *= corr[1];
    }

    // O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 8; ++dp) {
        uint32_t b0, b1, b2, b3;
        const int mi = lane >> 3;
        ldsm_x4_t(vb + swz(ks * 16 + (mi & 1) * 8 + (lane & 7), dp * 2 + (mi >> 1)), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pf[ks], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

  // finalise: O / l, bf16, token-major store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int row = row_a + r * 8;
    if (row < p.Sq) {
      bf16* op = O + (long)row * p.ldo + 2 * tq;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        *reinterpret_cast<uint32_t*>(op + nt * 8) =
            pack_bf16x2(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Cross-attention PROBABILITIES for the lyric-alignment callers (handler/lyric_timestamp.py:78-91,
// lyric_score.py: decoder(..., output_attentions=True)).  With output_attentions the reference runs the cross
// layers through eager_attention_forward (turbo :349-350), whose bf16 execution rounds at three points that
// this kernel reproduces: scores = bf16(q . k), bf16(scores * scaling), softmax in fp32, bf16(p).
// One CTA = 16 query rows of one (batch, head); K tiles of 64 keys are staged in smem once per 16 rows.
// The twice-rounded scores ARE bf16 values, so phase 1 parks them in the output rows themselves and phase 2
// (one warp per row) turns each row into probabilities in place: no smem limit on E.  SIMT FMA on purpose:
// this is a once-per-song side path (3 GFLOP per layer at 60 s), not part of the denoising loop.
constexpr int CP_ROWS = 16, CP_KEYS = 64;

__global__ void __launch_bounds__(256)
cross_probs_kernel(const bf16* __restrict__ q, long ldq, const bf16* __restrict__ k, long ldk,
                   bf16* __restrict__ out, int S, int E, int group, float scaling) {
  __shared__ __align__(16) float sq[CP_ROWS][128];
  __shared__ uint32_t sk[CP_KEYS][65];  // 64 bf16x2 words per key, +1: conflict-free when lane == key
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * CP_ROWS, h = blockIdx.y, b = blockIdx.z, H = gridDim.y, hk = h / group;
  for (int i = tid; i < CP_ROWS * 64; i += 256) {
    const int r = i >> 6, w = i & 63, row = r0 + r;
    float lo = 0.f, hi = 0.f;
    if (row < S)
      unpack_bf16x2(*reinterpret_cast<const uint32_t*>(q + ((long)b * S + row) * ldq + h * 128 + 2 * w), lo, hi);
    sq[r][2 * w] = lo;
    sq[r][2 * w + 1] = hi;
  }
  bf16* obase = out + (((long)b * H + h) * S + r0) * (long)E;
  const int key = tid & 63, rg = tid >> 6;  // this thread: one key of the tile x rows 4*rg .. 4*rg+3
  for (int e0 = 0; e0 < E; e0 += CP_KEYS) {
    __syncthreads();  // sq ready (first tile) / previous K tile consumed
    for (int i = tid; i < CP_KEYS * 64; i += 256) {
      const int kk = i >> 6, w = i & 63, e = e0 + kk;
      sk[kk][w] = e < E ? *reinterpret_cast<const uint32_t*>(k + ((long)b * E + e) * ldk + hk * 128 + 2 * w) : 0u;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int w = 0; w < 64; ++w) {
      float klo, khi;
      unpack_bf16x2(sk[key][w], klo, khi);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 qq = *reinterpret_cast<const float2*>(&sq[rg * 4 + j][2 * w]);
        acc[j] = fmaf(qq.x, klo, acc[j]);
        acc[j] = fmaf(qq.y, khi, acc[j]);
      }
    }
    const int e = e0 + key;
    if (e < E) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + rg * 4 + j < S) obase[(long)(rg * 4 + j) * E + e] = __float2bfloat16_rn(bf16_round(acc[j]) * scaling);
    }
  }
  __syncthreads();  // every score row of this CTA is written (and visible to the CTA)
  for (int r = warp; r < CP_ROWS && r0 + r < S; r += 8) {
    bf16* rowp = obase + (long)r * E;
    float m = -INFINITY;
    for (int e = lane; e < E; e += 32) m = fmaxf(m, __bfloat162float(rowp[e]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int e = lane; e < E; e += 32) sum += expf(__bfloat162float(rowp[e]) - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int e = lane; e < E; e += 32) rowp[e] = __float2bfloat16_rn(expf(__bfloat162float(rowp[e]) - m) / sum);
  }
}

}  // namespace

int launch_attention(const AttnParams& p, int heads, int batch, cudaStream_t stream) {
  static bool attr = false;
  const int smem = 5 * TILE_BYTES;
  if (!attr) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem));
    attr = true;
  }
  if (p.Sq <= 0 || p.Skv <= 0 || batch <= 0) return ACE_OK;
  dim3 grid(ceil_div(p.Sq, BM), heads, batch);
  // algorithmic work: 4 * Sq * (keys a query may see) * head_dim per head
  const double keys = p.window >= 0 ? (double)(p.Skv < 2 * p.window + 1 ? p.Skv : 2 * p.window + 1)
                                    : (double)p.Skv;
  const int kvh = heads / p.group;
  prof_begin(PROF_ATTN, 4.0 * p.Sq * keys * HD * heads * batch,
             2.0 * HD * batch * ((double)p.Sq * heads * 2 + (double)p.Skv * kvh * 2), stream);
  ACE_CUDA_CHECK(launch_kernel(attention_kernel, grid, dim3(ATT_THREADS), (size_t)smem, stream, p));
  prof_end(stream);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

int launch_cross_probs(const bf16* q, long ldq, const bf16* k, long ldk, bf16* out, int heads, int batch, int S,
                       int E, int group, cudaStream_t stream) {
  if (S <= 0 || E <= 0 || batch <= 0) return ACE_OK;
  ACE_REQUIRE(q && k && out, "cross_probs: null argument");
  ACE_REQUIRE((ldq % 2) == 0 && (ldk % 2) == 0, "cross_probs: odd row pitch");
  prof_begin(PROF_ATTN, 2.0 * S * E * HD * heads * batch, 2.0 * (double)S * E * heads * batch * 3, stream);
  ACE_CUDA_CHECK(launch_kernel(cross_probs_kernel, dim3(ceil_div(S, CP_ROWS), heads, batch), dim3(256), (size_t)0,
                               stream, q, ldq, k, ldk, out, S, E, group, 1.0f / sqrtf((float)HD)));
  prof_end(stream);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

}  // namespace ace
