// kernels.h — host-callable launchers of the non-GEMM kernels (attention, elementwise, codec).
#pragma once

#include "common.cuh"

namespace ace {

struct AttnParams {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* o;
  long ldq, ldk, ldv, ldo;  // row pitches in elements (token-major buffers)
  int Sq, Skv;              // per batch item
  int window;               // < 0: full attention, else |i - j| <= window
  int group;                // query heads per KV head
  float scale_log2;         // head_dim^-0.5 * log2(e)
  const int* kv_len;        // device [batch] or null: keys >= kv_len[b] are padding and masked (the
                            // padding-aware create_4d_mask of the condition encoders, turbo :53-132)
};
// softmax(q k^T / sqrt(128)) with the eager path's bf16 roundings -> out [batch][heads][S][E] (bf16); q rows
// [batch*S] at pitch ldq (head h at column 128 h), k rows [batch*E] at pitch ldk (kv head h / group)
int launch_cross_probs(const bf16* q, long ldq, const bf16* k, long ldk, bf16* out, int heads, int batch, int S,
                       int E, int group, cudaStream_t stream);

// tcgen05 / TMEM kernel (attention_tc.cu): tensor maps are encoded once per bound shape
struct AttnPlan {
  CUtensorMap tm_q, tm_k, tm_v, tm_o;  // tm_o: the output, stored by TMA from the (then idle) Q tile
  AttnParams p;
  int heads, batch;
};
int make_attn_plan(AttnPlan* plan, const AttnParams& p, int heads, int batch);
int launch_attention_tc(const AttnPlan& plan, cudaStream_t stream);
#ifdef ACE_PROBE
// probe builds: 1 = P through tensor memory (default), 0 = P through shared memory, -1 = environment default
void set_attention_p_in_tmem(int mode);
#endif

// x[b, t, :] = [ctx[b, t, 0:128] | xt[b, t, 0:64]] for t < T, zeros for T <= t < Tpad
int launch_concat_patches(const bf16* ctx, const bf16* xt, bf16* out, int B, int T, int Tpad,
                          cudaStream_t stream);

// out[r, :] = bf16( rmsnorm(h[r, :]) * w ) [* scale1p[b] + shift[b]]  (AdaLN, :490-496, 1488-1493)
//   shift / scale1p point at this op's rows of the per-step modulation table built by
//   the caller (batch stride mod_ld); scale1p == null -> plain RMSNorm.  (Used by the condition encoders: the DiT
//   step itself has no stand-alone norm kernel any more, see epilogues.cuh NormOut / NormIn.)
int launch_adaln_rmsnorm(const bf16* h, const bf16* w, const bf16* shift, const bf16* scale1p, long mod_ld,
                         bf16* out, int rows, int D, int rows_per_batch, float eps, cudaStream_t stream);

// Timestep-cache tables for `nb` consecutive entries starting at `entries` (see tcache_tables_kernel).
struct TCacheTablesArgs {
  uint8_t* entries;        // first entry
  size_t entry_bytes;
  size_t off_mods, off_outmod, off_cv, off_cout;  // byte offsets inside an entry
  int nb, L, D;
  const bf16* tables;      // [L][6][D] scale_shift_table of every layer
  const bf16* out_table;   // [2][D]
  const bf16* tproj;       // [nb][6 D]
  const bf16* temb;        // [nb][D]
  const bf16* self_norm0;  // layer 0's self_attn_norm / mlp_norm weights; layer l at + l * layer_stride
  const bf16* mlp_norm0;
  long layer_stride;
  const bf16* norm_out_w;  // [D]
};
int launch_tcache_tables(const TCacheTablesArgs& a, cudaStream_t stream);

// Timestep embedding (TimestepEmbedding.forward): sinusoid -> linear_1 -> SiLU -> linear_2 -> temb;
// SiLU -> time_proj -> proj.  Weights bf16 row-major [out, in].
struct TimeEmbedWeights {
  const bf16 *w1, *b1, *w2, *b2, *wp, *bp;
};
int launch_time_embed(const TimeEmbedWeights& w, const float* t /*[B] device*/, int B, int D,
                      bf16* scratch /*[B, 256 + 2D]*/, bf16* temb /*[B, D]*/, bf16* tproj /*[B, 6D]*/,
                      const bf16* add_temb, const bf16* add_proj, cudaStream_t stream);

// RoPE tables: cos/sin [S, 64] bf16 (fp32 math, then rounded — Qwen3RotaryEmbedding)
int launch_rope_tables(bf16* cos_tab, bf16* sin_tab, int S, float theta, cudaStream_t stream);

// ---- sampler-side elementwise ------------------------------------------------------------------
// xt <- bf16(xt - bf16(vt * dt))       (Euler, base :1977-1979 / turbo :1985-1991, :1975-1977)
int launch_euler(bf16* xt, const bf16* vt, float dt, long n, cudaStream_t stream, bf16* dup = nullptr);
// xt <- bf16(t_next * eps + (1 - t_next) * bf16(xt - bf16(vt * t_cur)))   (SDE re-noise)
int launch_sde(bf16* xt, const bf16* vt, const bf16* eps, float t_cur, float t_next, long n,
               cudaStream_t stream);
// APG (apg_guidance.py:33-56, dims=[1]): vt_out = cond + (scale-1) * orth(clip(momentum(diff)))
int launch_apg(const bf16* cond, const bf16* uncond, bf16* momentum /*[B,T,64] running average*/,
               int first_update, float momentum_coef, float norm_threshold, float guidance_scale,
               bf16* vt_out, int B, int T, cudaStream_t stream);
// ADG (apg_guidance.py:107-180) per (b, t) frame over the 64 channels
int launch_adg(const bf16* xt, const bf16* cond, const bf16* uncond, float sigma,
               float guidance_scale, float angle_clip, bf16* vt_out, int B, int T,
               cudaStream_t stream);
// Output path: per-sample |x| peak into peak[batch] (fp32) and in-place x / max(peak, 1), then (target_amp > 0)
// the front-end's normalisation to target_amp; latent sanity flags
// {any NaN/Inf, any non-zero} into flags[2].
int num_sms();  // runtime.cu
int launch_peak_normalize(float* wav, int batch, size_t n, float* peak, float target_amp, cudaStream_t stream);
int launch_latent_guard(const uint16_t* lat, size_t n, int* flags, cudaStream_t stream);

}  // namespace ace
