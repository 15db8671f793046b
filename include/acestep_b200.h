/* acestep_b200.h — C ABI of libacestep_b200.so (B200 / sm_100a only).
 *
 * Drop-in boundary for the ONE hot path of ACE-Step 1.5: the diffusion-transformer denoising loop
 * and the Oobleck VAE decode/encode.  The reference is 100 % Python and has no FFI of its own; the
 * entry points below are what its backend seam binds (the seam the MLX backend already uses):
 *
 *   DiT   : AceStepHandler._mlx_run_diffusion            acestep/core/generation/handler/diffusion.py:18-140
 *           selected in _execute_service_generate_diffusion   handler/service_generate_execute.py:144-194
 *           replaces model.generate_audio                 models/turbo/modeling_acestep_v15_turbo.py:1780-2001
 *                                                         models/base/modeling_acestep_v15_base.py:1783-1989
 *   codec : AceStepHandler.tiled_decode / _mlx_vae_decode handler/vae_decode.py:16-48, mlx_vae_decode_native.py:31-76
 *           AceStepHandler.tiled_encode / _mlx_vae_encode_sample   handler/vae_encode.py:15-43
 *
 * Conventions
 *   - every function returns 0 on success, a positive AceStatus otherwise; ace_last_error() gives
 *     the message of the calling thread's last failure.  No exceptions, no exit().
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers.  bf16 is passed as uint16_t.
 *   - the caller owns inputs, outputs and the workspace; a handle owns only its packed weights,
 *     the cross-attention K/V cache (inside the bound workspace) and its CUDA graphs.
 *   - all work is enqueued on the cudaStream_t passed as `void* stream` (0 = default stream);
 *     nothing synchronises unless documented.  A handle is single-caller (like the reference's
 *     handler, acestep_v15_pipeline.py:395-399) but independent handles may run concurrently.
 *   - there is NO CPU fallback: every entry point fails with ACE_ERR_CUDA on a non-sm_100 device.
 */
#ifndef ACESTEP_B200_H_
#define ACESTEP_B200_H_

#include <stddef.h>
#include <stdint.h>

/* The library is built with -fvisibility=hidden: only the entry points declared here are exported. */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef enum AceStatus {
  ACE_STATUS_OK = 0,
  ACE_STATUS_INVALID = 1,
  ACE_STATUS_CUDA = 2,
  ACE_STATUS_NOMEM = 3,
  ACE_STATUS_UNSUPPORTED = 4
} AceStatus;

const char* ace_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int ace_abi_version(void);
/* Checks that device `device` is compute capability 10.x and makes it current. */
int ace_init(int device);

/* ------------------------------------------------------------------------------------------ */
/* DiT (AceStepDiTModel, turbo modeling :1237-1504; config configuration_acestep_v15.py:148-260) */
/* ------------------------------------------------------------------------------------------ */
typedef struct AceDitConfig {
  int hidden_size;          /* 2048; multiple of 256 */
  int intermediate_size;    /* 6144; multiple of 64 */
  int num_layers;           /* 24 */
  int num_heads;            /* 16 */
  int num_kv_heads;         /* 8 */
  int head_dim;             /* must be 128 */
  int sliding_window;       /* 128 */
  int layer_is_sliding[64]; /* per layer: 1 = +-window band, 0 = full attention */
  float rope_theta;         /* 1e6 */
  float rms_eps;            /* 1e-6 */
} AceDitConfig;

typedef struct AceDit AceDit;

/* Number of bf16 elements ace_dit_create expects in `weights` (layout: DESIGN.md "Packed DiT weights",
 * produced by acestep_b200.pack.pack_dit from decoder.state_dict(), cf. models/mlx/dit_convert.py:33-66). */
size_t ace_dit_packed_elems(const AceDitConfig* cfg);
/* `weights` may be a host or a device pointer; it is copied. */
int ace_dit_create(AceDit** out, const AceDitConfig* cfg, const uint16_t* weights, size_t n_elems);
void ace_dit_destroy(AceDit* dit);

/* Workspace for an effective batch `bc` (songs x (2 if CFG)), `t` latent frames, `e` condition tokens. */
size_t ace_dit_workspace_bytes(const AceDit* dit, int bc, int t, int e);
/* Binds shapes + workspace: encodes all TMA descriptors and captures the denoising step into a CUDA
 * graph on first use.  Re-bind when bc / t / e or the workspace change. */
int ace_dit_bind(AceDit* dit, int bc, int t, int e, void* d_workspace, size_t workspace_bytes);

/* condition_embedder + cross-attention K/V of all layers for d_enc [bc, e, hidden] (bf16); replaces
 * the EncoderDecoderCache fill of the first decoder call (turbo modeling :307-330, 1356).
 * With CFG the caller passes rows [cond ; null_condition_emb] (base modeling :1905-1911). */
int ace_dit_set_condition(AceDit* dit, const uint16_t* d_enc, void* stream);

/* One velocity prediction: d_xt [bc,t,64], d_ctx [bc,t,128] (bf16), h_t [bc] (host floats, already
 * rounded to the model dtype by the caller) -> d_vt [bc,t,64] (bf16). */
int ace_dit_step(AceDit* dit, const uint16_t* d_xt, const uint16_t* d_ctx, const float* h_t,
                 uint16_t* d_vt, void* stream);

/* Everything of the forward that depends on the timestep only (TimestepEmbedding.forward :245-251, the six
 * modulation vectors of every layer :490-496, the output norm's :1488-1493, and the AdaLN shift terms folded into
 * bias rows of the consuming GEMMs) is kept in a per-handle cache keyed by the VALUE of t: ace_dit_step computes an
 * entry the first time it sees a timestep (about a millisecond) and finds it afterwards — a sampler visits the same
 * 8 / 27 / 60 values for every song.  This call fills the entries of a whole schedule in one batched pass
 * (h_t: n host floats, rounded like ace_dit_step's; duplicates and cached values are skipped).  Optional. */
int ace_dit_prepare_timesteps(AceDit* dit, const float* h_t, int n, void* stream);

/* One forward through layers [0, n_layers) that exports the CROSS-ATTENTION PROBABILITIES of each of them,
 * d_probs [n_layers][bc][heads][S][e] bf16 with S = ceil(t / 2) tokens, and stops there — what
 * `decoder(..., output_attentions=True, custom_layers_config=cfg, enable_early_exit=True)[2]` feeds the lyric
 * aligner (handler/lyric_timestamp.py:78-103, lyric_score.py; turbo :349-350, :1448-1482), with n_layers =
 * max(cfg) + 1.  Same roundings as the reference's eager path in bf16.  No velocity is produced. */
int ace_dit_cross_attentions(AceDit* dit, const uint16_t* d_xt, const uint16_t* d_ctx, const float* h_t,
                             int n_layers, uint16_t* d_probs, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Sampler update kernels (base :1945-1979, turbo :1975-1991, apg_guidance.py)                  */
/* ------------------------------------------------------------------------------------------ */
/* xt <- xt - vt * dt over n bf16 elements (n % 8 == 0) */
int ace_euler_step(uint16_t* d_xt, const uint16_t* d_vt, float dt, size_t n, void* stream);
/* Same update, and the new state is also written to `d_dup` (n elements): the unconditional half of the
 * next step's CFG batch, so the sampler needs no separate copy kernel between steps. */
int ace_euler_step_dup(uint16_t* d_xt, const uint16_t* d_vt, float dt, size_t n, uint16_t* d_dup, void* stream);
/* xt <- t_next * eps + (1 - t_next) * (xt - vt * t_cur) */
int ace_sde_step(uint16_t* d_xt, const uint16_t* d_vt, const uint16_t* d_eps, float t_cur,
                 float t_next, size_t n, void* stream);
/* APG over [b, t, 64]: d_momentum is the persistent running average (bf16, same shape);
 * first_update != 0 on the first guided step of a run. */
int ace_apg(const uint16_t* d_cond, const uint16_t* d_uncond, uint16_t* d_momentum, int first_update,
            float momentum, float norm_threshold, float guidance_scale, uint16_t* d_out, int b, int t,
            void* stream);
/* ADG over [b, t, 64] with per-frame angles. */
int ace_adg(const uint16_t* d_xt, const uint16_t* d_cond, const uint16_t* d_uncond, float sigma,
            float guidance_scale, float angle_clip, uint16_t* d_out, int b, int t, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Output path (handler/generate_music_decode.py:66-77 and :191-195)                            */
/* ------------------------------------------------------------------------------------------ */
/* Per-sample peak normalisation of decoded waveforms, in place: d_wav is [batch][n] fp32 (n = channels *
 * samples), d_peak [batch] fp32 receives max|x| of each sample, and a sample whose peak exceeds 1 is divided
 * by it (IEEE division, identical to `wav / peak.clamp(min=1)`; dividing the other samples by 1 is the
 * identity, so the reference's batch-wide `if torch.any(peak > 1)` needs no host decision).  A NaN in a
 * sample makes its peak NaN and leaves the sample untouched. */
int ace_peak_normalize(float* d_wav, int batch, size_t n, float* d_peak, void* stream);
/* The same pass followed, in the same kernel, by the front-end's `normalize_audio` (acestep/audio_utils.py:24-62;
 * on by default at -1 dB, inference.py:674-679) applied to the result: target_amp = 10^(dB/20) > 0; a sample whose
 * peak after the first stage is below 1e-6 is left as it is.  Bit-identical to the two torch passes; d_peak still
 * receives the raw peaks. */
int ace_peak_normalize_db(float* d_wav, int batch, size_t n, float* d_peak, float target_amp, void* stream);
/* Latent sanity guard over n bf16 elements: d_flags[0] = 1 if any NaN/Inf, d_flags[1] = 1 if any non-zero
 * (the reference raises on NaN/Inf and on all-zero latents before decoding). */
int ace_latent_guard(const uint16_t* d_lat, size_t n, int* d_flags, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Oobleck VAE (diffusers.AutoencoderOobleck; structure acestep/models/mlx/vae_model.py:149-230) */
/* ------------------------------------------------------------------------------------------ */
typedef struct AceVaeConfig {
  int num_stages;           /* 5 */
  int ratios[8];            /* encoder downsampling ratios, e.g. {2,4,4,6,10}; decoder uses reverse */
  int channel_multiples[8]; /* {1,2,4,8,16} */
  int encoder_hidden;       /* 128 */
  int decoder_channels;     /* 128 */
  int latent_channels;      /* 64 */
  int audio_channels;       /* 2 */
} AceVaeConfig;

typedef struct AceVae AceVae;

size_t ace_vae_packed_bytes(const AceVaeConfig* cfg);
/* `weights`: packed blob from acestep_b200.pack.pack_vae (weight-norm folded, tap-major bf16 conv
 * matrices, fp32 bias / Snake tables); host or device pointer; copied. */
int ace_vae_create(AceVae** out, const AceVaeConfig* cfg, const void* weights, size_t n_bytes);
void ace_vae_destroy(AceVae* vae);

size_t ace_vae_decode_workspace_bytes(const AceVae* vae, int frames);
size_t ace_vae_encode_workspace_bytes(const AceVae* vae, int samples);
/* d_z [frames, 64] bf16 (channels-last == the sampler's [T,64] layout) -> d_wav [2, frames*hop] fp32. */
int ace_vae_decode(AceVae* vae, const uint16_t* d_z, int frames, float* d_wav, void* d_workspace,
                   size_t workspace_bytes, void* stream);
/* d_wav [2, samples] fp32 (samples % hop == 0), d_eps [samples/hop, 64] bf16 posterior noise (may be
 * NULL: return the mean) -> d_z [samples/hop, 64] bf16 = mean + (softplus(scale) + 1e-4) * eps. */
int ace_vae_encode(AceVae* vae, const float* d_wav, int samples, const uint16_t* d_eps, uint16_t* d_z,
                   void* d_workspace, size_t workspace_bytes, void* stream);
/* The two halves of ace_vae_encode, for callers that encode the SAME audio repeatedly (reference / source audio of
 * successive requests: handler/conditioning_embed.py:18-69): the encoder's posterior moments are a deterministic
 * function of the audio, only the noise is fresh per call (`latent_dist.sample()`).
 *   ace_vae_encode_moments  : d_wav -> d_moments [samples/hop, 128] bf16 (mean | scale per frame)
 *   ace_vae_posterior_sample: d_moments, d_eps [frames, 64] (NULL: the mean) -> d_z [frames, 64], bit-identical to
 *                             what ace_vae_encode returns for the same audio and noise. */
int ace_vae_encode_moments(AceVae* vae, const float* d_wav, int samples, uint16_t* d_moments, void* d_workspace,
                           size_t workspace_bytes, void* stream);
int ace_vae_posterior_sample(const AceVae* vae, const uint16_t* d_moments, const uint16_t* d_eps, uint16_t* d_z,
                             int frames, void* stream);

/* Device addresses of the handle's static I/O slots inside the bound workspace (xt [bc,t,64],
 * ctx [bc,t,128], vt [bc,t,64]).  Passing these to ace_dit_step skips the staging copies, so a
 * sampler that keeps its state there runs the whole step from one CUDA graph launch. */
int ace_dit_io_slots(AceDit* dit, uint16_t** d_xt, uint16_t** d_ctx, uint16_t** d_vt);

/* ------------------------------------------------------------------------------------------ */
/* Condition encoders: the lyric / timbre transformer stacks of                */
/* AceStepConditionEncoder.forward, turbo modeling :1506-1552 (AceStepLyricEncoder :574-728,      */
/* AceStepTimbreEncoder :994-1175, AceStepEncoderLayer :371-437).  One handle = one stack:        */
/* embed_tokens Linear(in_dim -> hidden, bias) -> num_layers encoder layers -> final RMSNorm.      */
/* ------------------------------------------------------------------------------------------ */
typedef struct AceEncConfig {
  int hidden_size;          /* 2048; multiple of 256 */
  int intermediate_size;    /* 6144; multiple of 64 */
  int num_layers;           /* 8 (lyric) / 4 (timbre) */
  int num_heads;            /* 16 */
  int num_kv_heads;         /* 8 */
  int head_dim;             /* must be 128 */
  int sliding_window;       /* 128 */
  int layer_is_sliding[64]; /* per layer: 1 = +-window band, 0 = full attention */
  int in_dim;               /* embed_tokens input width: 1024 (lyric) / 64 (timbre); multiple of 64 */
  float rope_theta;         /* 1e6 */
  float rms_eps;            /* 1e-6 */
} AceEncConfig;
typedef struct AceEnc AceEnc;

/* Element count of the packed bf16 blob (acestep_b200/pack.py:pack_encoder is the one producer). */
size_t ace_enc_packed_elems(const AceEncConfig* cfg);
int ace_enc_create(AceEnc** out, const AceEncConfig* cfg, const uint16_t* weights, size_t n_elems);
void ace_enc_destroy(AceEnc* enc);
size_t ace_enc_workspace_bytes(const AceEnc* enc, int batch, int seq);
/* d_in [batch, seq, in_dim] bf16 -> d_out [batch, seq, hidden] bf16.  d_kv_len: DEVICE int[batch] of
 * valid (non-padding, left-aligned) tokens per sample, or NULL for no key-padding mask; semantics =
 * the reference's additive create_4d_mask (:53-132), including the uniform softmax of rows whose
 * whole band is padding. */
int ace_enc_forward(AceEnc* enc, const uint16_t* d_in, const int* d_kv_len, uint16_t* d_out, int batch, int seq,
                    void* ws, size_t ws_bytes, void* stream);
/* out[m, n] = bf16(a[m, k] . w[n, k]^T + bias[n]) on the tcgen05 GEMM; bias may be NULL
 * (text_projector, :1518).  k must be a multiple of 64. */
int ace_linear(const uint16_t* d_a, const uint16_t* d_w, const uint16_t* d_bias, uint16_t* d_out, int m, int n,
               int k, void* stream);

/* Residual finite-scalar quantizer of the audio tokenizer (AceStepAudioTokenizer.quantizer, turbo :1190-1194; the
 * class is the third-party vector_quantize_pytorch.ResidualFSQ, restated in oracle/tokenizer.py):
 *   d_x [m, dim] bf16 -> project_in (d_w_in [n_levels, dim], d_b_in [n_levels]) -> num_quantizers FSQ rounds in fp32
 *   with `levels` (host ints) -> project_out (d_w_out [dim, n_levels], d_b_out [dim]) -> d_q [m, dim] bf16,
 *   d_indices [m, num_quantizers] int32 (codebook index per round). */
int ace_fsq(const uint16_t* d_x, const uint16_t* d_w_in, const uint16_t* d_b_in, const int* levels, int n_levels,
            int num_quantizers, const uint16_t* d_w_out, const uint16_t* d_b_out, uint16_t* d_q, int* d_indices, int m,
            int dim, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Measurement hooks (bench.py): launch counter and per-launch CUDA-event profiling              */
/* ------------------------------------------------------------------------------------------ */
/* Number of kernels this library has launched in the calling process (graph replays included). */
uint64_t ace_launch_count(void);
/* Between start and stop every launch is bracketed by CUDA events on its own stream (CUDA graphs
 * are bypassed).  stop() synchronises and fills 4-element arrays indexed by category
 * {0: tcgen05 GEMM, 1: attention, 2: elementwise, 3: SIMT conv}: summed milliseconds, algorithmic
 * FLOPs, algorithmic HBM bytes and launch counts. */
void ace_profile_start(void);
int ace_profile_stop(float* ms, double* flops, double* bytes, int* launches);
/* After ace_profile_stop: the GEMM problem shapes of the window, sorted by total time (returns the
 * number written, at most max_out): per shape (M, N, K) the launch count and summed milliseconds. */
int ace_profile_gemm_shapes(int max_out, int* m, int* n, int* k, int* launches, float* ms);

/* ------------------------------------------------------------------------------------------ */
/* Single-op entry points (kernel-level parity tests and callers that need one op)              */
/* ------------------------------------------------------------------------------------------ */
/* softmax(q k^T / sqrt(128) [|i - j| <= window]) v on token-major bf16 buffers through the tcgen05 attention
 * kernel: q [batch*sq, heads*128], k / v [batch*skv, kv_heads*128], o [batch*sq, heads*128]; window < 0 = full.
 * (What AceStepAttention.forward hands to SDPA, modeling_acestep_v15_turbo.py:348-364.)  `ace_linear` above is the
 * matching single nn.Linear. */
int ace_attention(const uint16_t* d_q, const uint16_t* d_k, const uint16_t* d_v, uint16_t* d_o,
                  int batch, int heads, int kv_heads, int sq, int skv, int window, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Probe hooks: exported ONLY by libacestep_b200_probe.so (built with -DACE_PROBE for tools/ and  */
/* the two-path equivalence tests).  The release library has no A/B switch, no environment read, */
/* no alternative kernel.                                                                       */
/* ------------------------------------------------------------------------------------------ */
#ifdef ACE_PROBE
/* 1: route every GEMM through the scalar reference kernels (validates epilogues independently). */
void ace_debug_set_gemm_reference(int on);
/* Codec residual units of the 128-channel stages: 1 = single fused kernel (csrc/resunit.cuh),
 * 0 = two launches of the tap-shifted GEMM, -1 = default (fused unless ACE_VAE_FUSED=0).  Both paths
 * round at the same points, so their outputs are bit-identical (the test that uses this hook). */
void ace_debug_set_vae_fused(int on);
/* Attention: 1 = P goes back to tensor memory and P.V reads its A operand from there (default), 0 = P through
 * shared memory, -1 = default.  The stress test compares the two on long KV loops. */
void ace_debug_set_attention_p_in_tmem(int mode);
#endif

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* ACESTEP_B200_H_ */
