// cond.cu — the condition encoders (lyric / timbre transformer stacks) on the DiT's kernels.
//
// AceStepConditionEncoder.forward (modeling_acestep_v15_turbo.py:1506-1552) runs
// twice per request in the reference and is built from the same layer classes as the DiT:
//   AceStepLyricEncoder  (:574-728)   embed_tokens Linear(1024 -> 2048) + 8 x AceStepEncoderLayer + RMSNorm
//   AceStepTimbreEncoder (:994-1175)  embed_tokens Linear(64 -> 2048)   + 4 x AceStepEncoderLayer + RMSNorm
//   AceStepEncoderLayer  (:371-437)   RMSNorm -> self-attention (q/k RMSNorm, RoPE, GQA; even layers +-128
//                                     band, odd layers full; KEY-PADDING MASK APPLIED, unlike the DiT)
//                                     -> residual; RMSNorm -> SwiGLU MLP -> residual
// One AceEnc handle = one such stack.  Per layer 7 launches, all existing kernels:
//   rmsnorm -> [qkv GEMM + q/k RMSNorm + RoPE] -> flash attention(kv_len) -> [o GEMM + residual]
//   rmsnorm -> [gate|up GEMM + SwiGLU] -> [down GEMM + residual]
// text_projector (:1518) is a bias-free Linear: ace_linear.  pack_sequences / unpack_timbre_embeddings
// are index plumbing and stay in the PyTorch host code (acestep_b200/cond.py).
//
// The audio tokenizer's AttentionPooler (:730-856) and the AudioTokenDetokenizer (:859-990) are the same stack with
// the special token(s) joined AFTER embed_tokens, so they use a handle with in_dim = 0: the input is already
// [batch, seq, hidden] (acestep_b200/tokenizer.py builds it), sequences are pool_window_size (+1) tokens long.
// ace_fsq is the quantizer between them (ResidualFSQ of the third-party vector_quantize_pytorch, restated).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/acestep_b200.h"
#include "common.cuh"
#include "epilogues.cuh"
#include "gemm.cuh"
#include "kernels.h"

namespace ace {
namespace {

struct EncLayerW {
  const bf16 *in_norm, *post_norm, *qkv, *qn, *kn, *o, *gate_up, *down;
};

// Rows whose whole +-window band lies in the padding have every score replaced by finfo.min in the
// reference's additive mask (create_4d_mask :53-132), so their softmax is UNIFORM over all Skv keys
// (band-outside and padding keys included) and the output is bf16(1/Skv) * sum_j v_j.  The flash
// kernel writes zeros for such rows; this pass rewrites them.  They are padded positions, but the DiT
// cross-attends to padded condition tokens (its masks are dropped, :1381), so their values matter.
__global__ void __launch_bounds__(128)
masked_rows_uniform_kernel(const bf16* __restrict__ v, long ldv, bf16* __restrict__ o, long ldo,
                           const int* __restrict__ kv_len, int S, int window, int group) {
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;  // one thread per head_dim element
  const int first = kv_len[b] + window;                        // rows i >= first see no valid key in band
  if (first >= S) return;
  const bf16* vp = v + ((long)b * S) * ldv + (long)(h / group) * 128 + d;
  float sum = 0.f;
  for (int j = 0; j < S; ++j) sum += __bfloat162float(vp[(long)j * ldv]);
  const float val = bf16_round(bf16_round(1.0f / (float)S) * sum);
  for (int i = first; i < S; ++i) o[((long)b * S + i) * ldo + (long)h * 128 + d] = __float2bfloat16_rn(val);
}

// Residual FSQ (vector_quantize_pytorch ResidualFSQ / FSQ, restated — see oracle/tokenizer.py): one CTA per token.
//   y = bf16(W_in x + b_in)  [C <= 8 values];  per quantizer q, in fp32 (the library forces fp32 for the FSQ math):
//   z = residual / scale_q,  bounded = tanh(z + shift) * half_l - offset,  r = round-half-even(bounded),
//   code = bf16(r / (L // 2)) * scale_q,  index_q = sum_c (r_c + L_c // 2) * basis_c;  out = bf16(W_out sum_q code + b_out).
struct FsqLevels {
  int v[8];
};
__global__ void __launch_bounds__(256)
fsq_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w_in, const bf16* __restrict__ b_in, FsqLevels lv,
           int C, int nq, const bf16* __restrict__ w_out, const bf16* __restrict__ b_out, bf16* __restrict__ q_out,
           int* __restrict__ idx_out, int D) {
  pdl_trigger();
  pdl_wait();
  __shared__ float y[8];
  __shared__ float code_sum[8];
  __shared__ int idx_part[8][8];  // [quantizer][channel]
  const int m = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* xr = x + (size_t)m * D;
  for (int c = warp; c < C; c += 8) {
    float acc = 0.f;
    for (int k = lane * 2; k < D; k += 64) {
      float x0, x1, w0, w1;
      unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xr + k), x0, x1);
      unpack_bf16x2(*reinterpret_cast<const uint32_t*>(w_in + (size_t)c * D + k), w0, w1);
      acc = fmaf(x0, w0, acc);
      acc = fmaf(x1, w1, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[c] = bf16_round(acc + __bfloat162float(b_in[c]));
  }
  __syncthreads();
  if ((int)threadIdx.x < C) {
    const int c = threadIdx.x;
    const int L = lv.v[c];
    const float half_l = (float)(L - 1) * (1.0f + 1e-3f) / 2.0f;
    const float offset = (L % 2 == 0) ? 0.5f : 0.0f;
    const float shift = atanhf(offset / half_l);
    const float half_w = (float)(L / 2);
    int basis = 1;
    for (int i = 0; i < c; ++i) basis *= lv.v[i];
    float residual = y[c], out = 0.f;
    for (int qi = 0; qi < nq; ++qi) {
      const float scale = qi == 0 ? 1.0f : powf((float)(L - 1), -(float)qi);
      const float bounded = tanhf(residual / scale + shift) * half_l - offset;
      const float r = rintf(bounded);  // round half to even, like torch.round
      const float code = bf16_round(r / half_w) * scale;
      residual -= code;
      out += code;
      idx_part[qi][c] = ((int)r + L / 2) * basis;
    }
    code_sum[c] = bf16_round(out);
  }
  __syncthreads();
  if ((int)threadIdx.x < nq) {
    int s = 0;
    for (int c = 0; c < C; ++c) s += idx_part[threadIdx.x][c];
    idx_out[(size_t)m * nq + threadIdx.x] = s;
  }
  for (int d = threadIdx.x; d < D; d += 256) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(code_sum[c], __bfloat162float(w_out[(size_t)d * C + c]), acc);
    q_out[(size_t)m * D + d] = __float2bfloat16_rn(acc + __bfloat162float(b_out[d]));
  }
}

}  // namespace
}  // namespace ace

using namespace ace;

struct AceEnc {
  AceEncConfig cfg;
  int D, I, L, NQ, NKV, IN;
  bf16* weights = nullptr;
  size_t n_elems = 0;
  const bf16 *embed_w, *embed_b, *final_norm;
  std::vector<EncLayerW> lw;
};

extern "C" {

size_t ace_enc_packed_elems(const AceEncConfig* c) {
  if (!c) return 0;
  const size_t D = c->hidden_size, I = c->intermediate_size, L = c->num_layers, IN = c->in_dim;
  const size_t NQ = (size_t)c->num_heads * c->head_dim, NKV = (size_t)c->num_kv_heads * c->head_dim;
  const size_t per = 2 * D + (NQ + 2 * NKV) * D + 256 + D * NQ + 2 * I * D + D * I;
  return D * IN + (IN ? D : 0) + D + L * per;  // in_dim 0: no embed_tokens in the blob
}

int ace_enc_create(AceEnc** out, const AceEncConfig* cfg, const uint16_t* weights, size_t n_elems) {
  ACE_REQUIRE(out && cfg && weights, "ace_enc_create: null argument");
  {
    int dev = 0, major = 0;
    ACE_CUDA_CHECK(cudaGetDevice(&dev));
    ACE_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    ACE_REQUIRE(major == 10, "libacestep_b200 requires an sm_100 (B200) device, found %d.x", major);
  }
  ACE_REQUIRE(cfg->head_dim == 128, "head_dim %d unsupported (kernels are specialised for 128)", cfg->head_dim);
  ACE_REQUIRE(cfg->hidden_size % 256 == 0 && cfg->intermediate_size % 64 == 0 && cfg->in_dim % 64 == 0,
              "hidden_size must be a multiple of 256, intermediate_size and in_dim of 64");
  ACE_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= 64, "num_layers %d out of range", cfg->num_layers);
  ACE_REQUIRE(cfg->num_kv_heads >= 1 && cfg->num_heads % cfg->num_kv_heads == 0, "bad head counts");
  ACE_REQUIRE(n_elems == ace_enc_packed_elems(cfg), "packed encoder blob has %zu elements, expected %zu", n_elems,
              ace_enc_packed_elems(cfg));
  AceEnc* e = new AceEnc();
  e->cfg = *cfg;
  e->D = cfg->hidden_size;
  e->I = cfg->intermediate_size;
  e->L = cfg->num_layers;
  e->NQ = cfg->num_heads * 128;
  e->NKV = cfg->num_kv_heads * 128;
  e->IN = cfg->in_dim;
  e->n_elems = n_elems;
  if (cudaMalloc(&e->weights, n_elems * sizeof(bf16)) != cudaSuccess) {
    delete e;
    set_error("cudaMalloc of %zu encoder weight bytes failed", n_elems * sizeof(bf16));
    return ACE_ERR_NOMEM;
  }
  cudaError_t ce = cudaMemcpy(e->weights, weights, n_elems * sizeof(bf16), cudaMemcpyDefault);
  if (ce != cudaSuccess) {
    cudaFree(e->weights);
    delete e;
    set_error("encoder weight upload failed: %s", cudaGetErrorString(ce));
    return ACE_ERR_CUDA;
  }
  const size_t D = e->D, I = e->I, NQ = e->NQ, NKV = e->NKV;
  const bf16* p = e->weights;
  auto take = [&](size_t n) {
    const bf16* r = p;
    p += n;
    return r;
  };
  e->embed_w = take(D * e->IN);
  e->embed_b = take(e->IN ? D : 0);
  e->final_norm = take(D);
  e->lw.resize(e->L);
  for (EncLayerW& w : e->lw) {
    w.in_norm = take(D);
    w.post_norm = take(D);
    w.qkv = take((NQ + 2 * NKV) * D);
    w.qn = take(128);
    w.kn = take(128);
    w.o = take(D * NQ);
    w.gate_up = take(2 * I * D);
    w.down = take(D * I);
  }
  *out = e;
  return ACE_OK;
}

void ace_enc_destroy(AceEnc* e) {
  if (!e) return;
  cudaFree(e->weights);
  delete e;
}

static size_t enc_align(size_t x) { return (x + 255) & ~(size_t)255; }

size_t ace_enc_workspace_bytes(const AceEnc* e, int batch, int seq) {
  if (!e || batch <= 0 || seq <= 0) return 0;
  const size_t M = (size_t)batch * seq;
  return enc_align(M * e->D * 2) * 2 + enc_align(M * (e->NQ + 2 * e->NKV) * 2) + enc_align(M * e->NQ * 2) +
         enc_align(M * e->I * 2) + 2 * enc_align((size_t)seq * 64 * 2) + 256;
}

int ace_enc_forward(AceEnc* e, const uint16_t* d_in, const int* d_kv_len, uint16_t* d_out, int batch, int seq,
                    void* ws, size_t ws_bytes, void* stream) {
  ACE_REQUIRE(e && d_in && d_out && ws, "ace_enc_forward: null argument");
  ACE_REQUIRE(batch >= 1 && seq >= 1, "bad shape batch=%d seq=%d", batch, seq);
  ACE_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)d_in & 15) == 0, "unaligned buffer");
  ACE_REQUIRE(ws_bytes >= ace_enc_workspace_bytes(e, batch, seq), "encoder workspace too small: %zu < %zu", ws_bytes,
              ace_enc_workspace_bytes(e, batch, seq));
  cudaStream_t st = (cudaStream_t)stream;
  const int D = e->D, I = e->I, NQ = e->NQ, NKV = e->NKV, S = seq, M = batch * seq;
  const long QKVW = NQ + 2 * NKV;
  uint8_t* wp = (uint8_t*)ws;
  auto carve = [&](size_t bytes) {
    uint8_t* r = wp;
    wp += enc_align(bytes);
    return reinterpret_cast<bf16*>(r);
  };
  bf16* h = carve((size_t)M * D * 2);
  bf16* hn = carve((size_t)M * D * 2);
  bf16* qkv = carve((size_t)M * QKVW * 2);
  bf16* attn = carve((size_t)M * NQ * 2);
  bf16* act = carve((size_t)M * I * 2);
  bf16* rope_cos = carve((size_t)S * 64 * 2);
  bf16* rope_sin = carve((size_t)S * 64 * 2);
  const float eps = e->cfg.rms_eps;
  const float scale_log2 = (1.0f / sqrtf(128.0f)) * 1.4426950408889634f;
  const int group = e->cfg.num_heads / e->cfg.num_kv_heads;

  ACE_PROPAGATE(launch_rope_tables(rope_cos, rope_sin, S, e->cfg.rope_theta, st));
  if (e->IN == 0) {  // pre-embedded input
    ACE_CUDA_CHECK(cudaMemcpyAsync(h, d_in, (size_t)M * D * 2, cudaMemcpyDeviceToDevice, st));
  } else {
    GemmPlan pe;
    ACE_PROPAGATE(make_gemm_plan(&pe, (const bf16*)d_in, M, e->IN, e->IN, e->embed_w, D, e->IN, M, 1, nullptr, 0));
    ACE_PROPAGATE(launch_gemm(pe, EpiBias{h, (long)D, e->embed_b}, st));
  }
  for (int l = 0; l < e->L; ++l) {
    const EncLayerW& w = e->lw[l];
    const int window = e->cfg.layer_is_sliding[l] ? e->cfg.sliding_window : -1;
    GemmPlan pq, po, pg, pd;
    ACE_PROPAGATE(make_gemm_plan(&pq, hn, M, D, D, w.qkv, NQ + 2 * NKV, D, M, 1, nullptr, 0));
    ACE_PROPAGATE(make_gemm_plan(&po, attn, M, NQ, NQ, w.o, D, NQ, M, 1, nullptr, -192));
    ACE_PROPAGATE(make_gemm_plan(&pg, hn, M, D, D, w.gate_up, 2 * I, D, M, 1, nullptr, 0));
    ACE_PROPAGATE(make_gemm_plan(&pd, act, M, I, I, w.down, D, I, M, 1, nullptr, -192));
    ACE_PROPAGATE(launch_adaln_rmsnorm(h, w.in_norm, nullptr, nullptr, 0, hn, M, D, S, eps, st));
    ACE_PROPAGATE(launch_gemm(pq, EpiQKV{qkv, QKVW, NQ, NKV, w.qn, w.kn, rope_cos, rope_sin, S, eps}, st));
    AttnParams ap{qkv, qkv + NQ, qkv + NQ + NKV, attn, QKVW, QKVW, QKVW, (long)NQ, S, S, window, group, scale_log2,
                  d_kv_len};
    {
      AttnPlan plan;
      ACE_PROPAGATE(make_attn_plan(&plan, ap, e->cfg.num_heads, batch));
      ACE_PROPAGATE(launch_attention_tc(plan, st));
    }
    if (d_kv_len != nullptr && window >= 0) {
      prof_begin(PROF_ELEM, 0.0, (double)batch * S * NKV * 2, st);
      ACE_CUDA_CHECK(launch_kernel(masked_rows_uniform_kernel, dim3(e->cfg.num_heads, batch), dim3(128), (size_t)0, st,
                                   (const bf16*)(qkv + NQ + NKV), QKVW, attn, (long)NQ, d_kv_len, S, window, group));
      prof_end(st);
    }
    ACE_PROPAGATE(launch_gemm(po, EpiGatedResid{h, (long)D, nullptr, 0, S}, st));
    ACE_PROPAGATE(launch_adaln_rmsnorm(h, w.post_norm, nullptr, nullptr, 0, hn, M, D, S, eps, st));
    ACE_PROPAGATE(launch_gemm(pg, EpiSwiGLU{act, (long)I}, st));
    ACE_PROPAGATE(launch_gemm(pd, EpiGatedResid{h, (long)D, nullptr, 0, S}, st));
  }
  ACE_PROPAGATE(launch_adaln_rmsnorm(h, e->final_norm, nullptr, nullptr, 0, (bf16*)d_out, M, D, S, eps, st));
  return ACE_OK;
}

int ace_fsq(const uint16_t* d_x, const uint16_t* d_w_in, const uint16_t* d_b_in, const int* levels, int n_levels,
            int num_quantizers, const uint16_t* d_w_out, const uint16_t* d_b_out, uint16_t* d_q, int* d_indices, int m,
            int dim, void* stream) {
  ACE_REQUIRE(d_x && d_w_in && d_b_in && levels && d_w_out && d_b_out && d_q && d_indices, "ace_fsq: null argument");
  ACE_REQUIRE(n_levels >= 1 && n_levels <= 8 && num_quantizers >= 1 && num_quantizers <= 8,
              "ace_fsq: %d levels / %d quantizers out of range [1, 8]", n_levels, num_quantizers);
  ACE_REQUIRE(m >= 1 && dim >= 64 && dim % 64 == 0, "ace_fsq: bad shape m=%d dim=%d", m, dim);
  FsqLevels lv;
  long codes = 1;
  for (int i = 0; i < 8; ++i) {
    lv.v[i] = i < n_levels ? levels[i] : 1;
    ACE_REQUIRE(lv.v[i] >= 1 && lv.v[i] <= 1024, "ace_fsq: level %d out of range", lv.v[i]);
    codes *= lv.v[i];
  }
  ACE_REQUIRE(codes < (1L << 31), "ace_fsq: codebook of %ld entries does not fit int32 indices", codes);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PROF_ELEM, 4.0 * m * dim * n_levels, 4.0 * m * dim, st);
  ACE_CUDA_CHECK(launch_kernel(fsq_kernel, dim3(m), dim3(256), (size_t)0, st, (const bf16*)d_x, (const bf16*)d_w_in,
                               (const bf16*)d_b_in, lv, n_levels, num_quantizers, (const bf16*)d_w_out,
                               (const bf16*)d_b_out, (bf16*)d_q, d_indices, dim));
  prof_end(st);
  return ACE_OK;
}

int ace_linear(const uint16_t* d_a, const uint16_t* d_w, const uint16_t* d_bias, uint16_t* d_out, int m, int n, int k,
               void* stream) {
  ACE_REQUIRE(d_a && d_w && d_out, "ace_linear: null argument");
  ACE_REQUIRE(m >= 1 && n >= 1 && k >= 64 && k % 64 == 0, "ace_linear: bad shape m=%d n=%d k=%d (k must be a multiple of 64)",
              m, n, k);
  GemmPlan p;
  ACE_PROPAGATE(make_gemm_plan(&p, (const bf16*)d_a, m, k, k, (const bf16*)d_w, n, k, m, 1, nullptr, 0));
  return launch_gemm(p, EpiBias{(bf16*)d_out, (long)n, (const bf16*)d_bias}, (cudaStream_t)stream);
}

}  // extern "C"
