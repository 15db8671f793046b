// attention.cu — cross-attention PROBABILITIES for the lyric-alignment callers (SIMT side path).
//
// The attention of the denoising loop itself (full / sliding-window / cross, modeling_acestep_v15_turbo.py:286-368)
// is the tcgen05 / TMEM kernel in attention_tc.cu.  (The round-1 mma.sync kernel that used to live here was
// retired once the tcgen05 kernel had replaced it everywhere; it is in the git history.)
#include "common.cuh"
#include "kernels.h"

namespace ace {

namespace {

constexpr int HD = 128;

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
