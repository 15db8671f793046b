// vae.cu — Oobleck VAE decode / encode on the tap-shifted tcgen05 GEMM (channels-last, bf16).
//
// Structure follows diffusers.AutoencoderOobleck as mirrored in-tree by
// acestep/models/mlx/vae_model.py:62-230 (residual unit, encoder/decoder blocks) with the
// weight-norm fold of mlx/vae_convert.py:18-34 done at pack time.
//
// Data layout: every activation is a [L, C] bf16 matrix (time-major, channels contiguous), which
// makes each convolution a GEMM over TMA tiles of that matrix:
//   Conv1d k=7 dilation d       : 7 taps, A row shift (tap-3)*d, K = 7*Cin      (zero pad = TMA OOB fill)
//   Conv1d k=1                  : 1 tap
//   ConvTranspose1d k=2s str. s : 2 taps {0,-1}, N = s*Cout, output row q lands at frame q*s - pad
//   Conv1d k=2s stride s        : 1 tap over an overlapping-row view (row pitch s*Cin, row length
//                                 2s*Cin) of the zero-haloed input
// Snake activations never run as separate passes: a producer's epilogue writes both x (for the
// residual) and snake(x) with the NEXT op's alpha/beta (the next conv's A operand).
// Only the 2-channel ends (encoder conv1 2->128, decoder conv2 128->2) are SIMT kernels.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/acestep_b200.h"
#include "common.cuh"
#include "epilogues.cuh"
#include "gemm.cuh"
#include "kernels.h"
#include "resunit.cuh"

namespace ace {

struct ConvW {
  const bf16* w = nullptr;    // [N, ntaps*Cin] tap-major
  const float* bias = nullptr;  // [Cout] or null
};
struct SnakeW {
  const float* a = nullptr;   // exp(alpha)           [C]
  const float* ib = nullptr;  // 1/(exp(beta)+1e-9)   [C]
};
struct ResUnitW {
  SnakeW s1, s2;
  ConvW c1, c2;
};
struct StageW {
  SnakeW snake;
  ConvW conv;  // decoder: conv_t1 [s*Cout, 2*Cin]; encoder: strided conv [Cout, 2s*Cin]
  ResUnitW ru[3];
  int cin, cout, stride, pad;
};

namespace {

constexpr int HALO = 16;  // zero rows kept before/after every activation buffer

// decoder conv2: Snake'd [L,C] bf16 -> planar fp32 [2, L]; k=7, pad 3, no bias.
// HBM-bound by construction (reads L*C*2 bytes, 10 GFLOP at C2): the work is SIMT FMA, blocked so
// that shared-memory traffic stays far below the FMA rate.  A block covers FC_POS output positions;
// a thread owns 4 consecutive positions x both output channels x 1/8 of the input channels (two
// 8-channel groups per 128 channels, interleaved so a warp's 16-byte reads are contiguous), i.e.
// 10 input rows and 14 weight vectors feed 4*2*7 = 56 dot products; the 8 channel lanes are then
// reduced with shuffles.
constexpr int FC_POS = 128;  // output positions per block (256 threads = 32 position groups x 8 channel lanes)
__global__ void __launch_bounds__(256)
final_conv_kernel(const bf16* __restrict__ x, const float* __restrict__ w /*[2][7][C]*/, float* __restrict__ out,
                  long L, int C) {
  extern __shared__ __align__(16) uint8_t fc_smem[];
  pdl_trigger();
  float* sw = reinterpret_cast<float*>(fc_smem);                       // [2][7][C] fp32
  uint4* sx = reinterpret_cast<uint4*>(fc_smem + (size_t)14 * C * 4);  // [(FC_POS + 6)][C] bf16
  const int tid = threadIdx.x;
  const long l0 = (long)blockIdx.x * FC_POS;
  for (int i = tid; i < 14 * C; i += 256) sw[i] = w[i];  // weights: not produced by the previous kernel
  pdl_wait();
  const int row_u4 = C / 8;  // uint4 per row
  for (int i = tid; i < (FC_POS + 6) * row_u4; i += 256) {
    const int r = i / row_u4;
    const long l = l0 - 3 + r;
    sx[i] = (l >= 0 && l < L) ? reinterpret_cast<const uint4*>(x + l * C)[i - r * row_u4] : make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const int cl = tid & 7, pg = tid >> 3;  // channel lane, position group
  float acc[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.f;
  for (int cg = cl; cg < row_u4; cg += 8) {  // this lane's 8-channel groups
    float xv[10][8];
#pragma unroll
    for (int rr = 0; rr < 10; ++rr) {
      const uint4 q = sx[(4 * pg + rr) * row_u4 + cg];
      unpack_bf16x2(q.x, xv[rr][0], xv[rr][1]);
      unpack_bf16x2(q.y, xv[rr][2], xv[rr][3]);
      unpack_bf16x2(q.z, xv[rr][4], xv[rr][5]);
      unpack_bf16x2(q.w, xv[rr][6], xv[rr][7]);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
#pragma unroll
      for (int co = 0; co < 2; ++co) {
        const float4 wa = *reinterpret_cast<const float4*>(sw + (co * 7 + k) * C + cg * 8);
        const float4 wb = *reinterpret_cast<const float4*>(sw + (co * 7 + k) * C + cg * 8 + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* xr = xv[j + k];
          float a = acc[j][co];
          a = fmaf(xr[0], wa.x, a); a = fmaf(xr[1], wa.y, a); a = fmaf(xr[2], wa.z, a); a = fmaf(xr[3], wa.w, a);
          a = fmaf(xr[4], wb.x, a); a = fmaf(xr[5], wb.y, a); a = fmaf(xr[6], wb.z, a); a = fmaf(xr[7], wb.w, a);
          acc[j][co] = a;
        }
      }
    }
  }
  // reduce over the 8 channel lanes (adjacent lanes of the warp); lane cl then stores value cl of the 8
  float mine = 0.f;
#pragma unroll
  for (int co = 0; co < 2; ++co)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[j][co];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (cl == co * 4 + j) mine = v;
    }
  const long l = l0 + 4 * pg + (cl & 3);
  if (l < L) out[(long)(cl >> 2) * L + l] = bf16_round(mine);
}

// encoder conv1: planar fp32 [2, N] -> [N, C] bf16 (+ Snake'd copy); k=7, pad 3, with bias.
// HBM-bound on its two [N, C] outputs (3 GB for 120 s of audio).  A thread owns 2 adjacent channels
// x FC1_S consecutive samples: its 28 weights and 2 x (FC1_S + 6) inputs live in registers, the input
// loads are warp-wide broadcasts, and each store instruction of a warp writes 128 contiguous bytes
// (32 lanes x one bf16 pair of the same sample).
constexpr int FC1_S = 8;
__global__ void __launch_bounds__(256)
first_conv_kernel(const float* __restrict__ wav, const float* __restrict__ w /*[C][2][7]*/,
                  const float* __restrict__ bias, const float* __restrict__ sn_a, const float* __restrict__ sn_ib,
                  bf16* __restrict__ out, bf16* __restrict__ out_snake, long N, int C) {
  pdl_trigger();
  const int pairs = C / 2;                       // channel pairs per sample
  const int groups_per_block = 256 / pairs;      // sample groups handled by one block (pairs divides 256)
  const int cp = threadIdx.x % pairs, sg = threadIdx.x / pairs;
  const long n0 = ((long)blockIdx.x * groups_per_block + sg) * FC1_S;
  const int c = 2 * cp;
  float wr[2][14];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int k = 0; k < 14; ++k) wr[j][k] = w[(c + j) * 14 + k];
  const float b0 = bias[c], b1 = bias[c + 1];
  const float a0 = sn_a[c], a1 = sn_a[c + 1], i0 = sn_ib[c], i1 = sn_ib[c + 1];
  pdl_wait();
  if (n0 >= N) return;
  float xin[2][FC1_S + 6];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch)
#pragma unroll
    for (int k = 0; k < FC1_S + 6; ++k) {
      const long m = n0 - 3 + k;
      xin[ch][k] = (m >= 0 && m < N) ? bf16_round(wav[ch * N + m]) : 0.f;
    }
#pragma unroll
  for (int s = 0; s < FC1_S; ++s) {
    if (n0 + s >= N) break;
    float acc0 = b0, acc1 = b1;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        acc0 = fmaf(xin[ch][s + k], wr[0][ch * 7 + k], acc0);
        acc1 = fmaf(xin[ch][s + k], wr[1][ch * 7 + k], acc1);
      }
    const float v0 = bf16_round(acc0), v1 = bf16_round(acc1);
    const long o = (n0 + s) * C + c;
    *reinterpret_cast<uint32_t*>(out + o) = pack_bf16x2(v0, v1);
    *reinterpret_cast<uint32_t*>(out_snake + o) = pack_bf16x2(snake_f(v0, a0, i0), snake_f(v1, a1, i1));
  }
}

// posterior: moments [L, 128] bf16 (mean | scale) -> z = mean + (softplus(scale) + 1e-4) * eps
__global__ void posterior_kernel(const bf16* __restrict__ mom, const bf16* __restrict__ eps, bf16* __restrict__ z,
                                 long L, int Cz) {
  pdl_trigger();
  pdl_wait();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * Cz) return;
  const long l = idx / Cz;
  const int c = idx % Cz;
  const float mean = __bfloat162float(mom[l * 2 * Cz + c]);
  float out = mean;
  if (eps) {
    const float sc = __bfloat162float(mom[l * 2 * Cz + Cz + c]);
    const float sp = sc > 20.f ? sc : log1pf(expf(sc));
    const float std_ = bf16_round(bf16_round(sp) + 1e-4f);
    out = bf16_round(mean + bf16_round(std_ * __bfloat162float(eps[idx])));
  }
  z[idx] = __float2bfloat16_rn(out);
}

struct Buf {
  bf16* p;  // first valid row (HALO zero rows sit before it)
};

}  // namespace
}  // namespace ace

using namespace ace;

struct AceVae {
  AceVaeConfig cfg;
  uint8_t* blob = nullptr;
  size_t blob_bytes = 0;
  int hop = 1;
  // decoder
  ConvW dec_conv1;
  std::vector<StageW> dec;
  SnakeW dec_snake;
  const float* dec_conv2 = nullptr;  // [2][7][C] fp32
  // encoder
  const float* enc_conv1_w = nullptr;  // [C][2][7] fp32
  const float* enc_conv1_b = nullptr;
  std::vector<StageW> enc;
  SnakeW enc_snake;
  ConvW enc_conv2;  // [2*Cz, 3*Cin]
};

namespace {

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Blob layout walker shared by sizing (base == nullptr) and binding.
struct BlobWalker {
  uint8_t* base;
  size_t off = 0;
  const bf16* bf(size_t n) {
    off = align256(off);
    const bf16* p = base ? reinterpret_cast<const bf16*>(base + off) : nullptr;
    off += n * 2;
    return p;
  }
  const float* f32(size_t n) {
    off = align256(off);
    const float* p = base ? reinterpret_cast<const float*>(base + off) : nullptr;
    off += n * 4;
    return p;
  }
};

void walk_conv(BlobWalker& b, ConvW& c, size_t n, size_t k, bool bias, size_t nbias) {
  c.w = b.bf(n * k);
  c.bias = bias ? b.f32(nbias) : nullptr;
}
void walk_snake(BlobWalker& b, SnakeW& s, size_t c) {
  s.a = b.f32(c);
  s.ib = b.f32(c);
}
void walk_ru(BlobWalker& b, ResUnitW& r, size_t c) {
  walk_snake(b, r.s1, c);
  walk_conv(b, r.c1, c, 7 * c, true, c);
  walk_snake(b, r.s2, c);
  walk_conv(b, r.c2, c, c, true, c);
}

size_t walk_blob(AceVae* v, uint8_t* base) {
  const AceVaeConfig& c = v->cfg;
  const int n = c.num_stages;
  BlobWalker b{base};
  int cm[9];
  cm[0] = 1;
  for (int i = 0; i < n; ++i) cm[i + 1] = c.channel_multiples[i];
  // ---- decoder ----
  const int C = c.decoder_channels;
  walk_conv(b, v->dec_conv1, (size_t)C * cm[n], 7 * (size_t)c.latent_channels, true, (size_t)C * cm[n]);
  v->dec.resize(n);
  for (int i = 0; i < n; ++i) {
    StageW& s = v->dec[i];
    s.stride = c.ratios[n - 1 - i];
    s.pad = (s.stride + 1) / 2;
    s.cin = C * cm[n - i];
    s.cout = C * cm[n - i - 1];
    walk_snake(b, s.snake, s.cin);
    walk_conv(b, s.conv, (size_t)s.stride * s.cout, 2 * (size_t)s.cin, true, s.cout);
    for (int j = 0; j < 3; ++j) walk_ru(b, s.ru[j], s.cout);
  }
  walk_snake(b, v->dec_snake, C);
  v->dec_conv2 = b.f32((size_t)c.audio_channels * 7 * C);
  // ---- encoder ----
  const int H = c.encoder_hidden;
  v->enc_conv1_w = b.f32((size_t)H * c.audio_channels * 7);
  v->enc_conv1_b = b.f32(H);
  v->enc.resize(n);
  for (int i = 0; i < n; ++i) {
    StageW& s = v->enc[i];
    s.stride = c.ratios[i];
    s.pad = (s.stride + 1) / 2;
    s.cin = H * cm[i];
    s.cout = H * cm[i + 1];
    for (int j = 0; j < 3; ++j) walk_ru(b, s.ru[j], s.cin);
    walk_snake(b, s.snake, s.cin);
    walk_conv(b, s.conv, s.cout, 2 * (size_t)s.stride * s.cin, true, s.cout);
  }
  walk_snake(b, v->enc_snake, (size_t)H * cm[n]);
  walk_conv(b, v->enc_conv2, H, 3 * (size_t)H * cm[n], true, H);
  return align256(b.off);
}

// conv as GEMM over [L, Cin] -> epilogue; `shifts` row shifts per tap.
int conv_gemm(const bf16* a, long a_rows, int kc, long a_ld, const ConvW& w, int n, long m, int ntaps,
              const int* shifts, const EpiConv& epi, cudaStream_t st) {
  GemmPlan p;
  ACE_PROPAGATE(make_gemm_plan(&p, a, (int)a_rows, kc, a_ld, w.w, n, (long)ntaps * kc, (int)m, ntaps, shifts, 0));
  return launch_gemm(p, epi, st);
}

// Residual unit on x [L, C] (and its Snake'd copy xs): returns new x / xs in the `o*` buffers.
//   h  = conv7_d(snake1(x))            A = xs           epilogue: snake2 -> hs
//   x' = x + conv1(snake2(h))          A = hs           epilogue: + x, also snake_next(x') -> oxs
#ifdef ACE_PROBE
int g_fused_override = -1;  // ace_debug_set_vae_fused: -1 = environment default, 0 / 1 = forced
bool fused_res_unit_enabled() {
  if (g_fused_override >= 0) return g_fused_override != 0;
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ACE_VAE_FUSED");  // ACE_VAE_FUSED=0: two-launch path (A/B measurements)
    on = !(e && e[0] == '0');
  }
  return on != 0;
}
#else
constexpr bool fused_res_unit_enabled() { return true; }
#endif

int res_unit(const ResUnitW& r, int dil, long L, int C, const bf16* x, const bf16* xs, bf16* hs, bf16* ox,
             bf16* oxs, const SnakeW& next, cudaStream_t st) {
  if (C == RU_C && fused_res_unit_enabled() && !gemm_debug_reference()) {
    // 128-channel stages (2/3 of the codec's HBM traffic): the whole unit in one kernel, hs stays on chip
    RuParams p{(int)L, dil, r.c1.bias, r.s2.a, r.s2.ib, r.c2.bias, next.a, next.ib, ox, oxs};
    return launch_res_unit_fused(xs, x, r.c1.w, r.c2.w, p, st);
  }
  int sh[7];
  for (int k = 0; k < 7; ++k) sh[k] = (k - 3) * dil;
  EpiConv e1{nullptr, hs, nullptr, r.c1.bias, r.s2.a, r.s2.ib, (long)C, 0, L * C, C};
  ACE_PROPAGATE(conv_gemm(xs, L, C, C, r.c1, C, L, 7, sh, e1, st));
  EpiConv e2{ox, oxs, x, r.c2.bias, next.a, next.ib, (long)C, 0, L * C, C};
  const int z = 0;
  return conv_gemm(hs, L, C, C, r.c2, C, L, 1, &z, e2, st);
}

struct DecodeLayout {
  size_t bytes;
  size_t buf_elems;  // per ping-pong buffer (incl. halo)
};

// Largest activation (elements) across decoder stages for `frames` latent frames.
size_t dec_max_elems(const AceVae* v, long frames) {
  size_t mx = (size_t)frames * v->dec[0].cin;
  long L = frames;
  for (const StageW& s : v->dec) {
    L = L * s.stride + ((s.stride & 1) ? -1 : 0);
    mx = mx > (size_t)L * s.cout ? mx : (size_t)L * s.cout;
  }
  return mx;
}
size_t enc_max_elems(const AceVae* v, long samples) {
  size_t mx = (size_t)samples * v->enc[0].cin;
  long L = samples;
  for (const StageW& s : v->enc) {
    mx = mx > (size_t)L * s.cin ? mx : (size_t)L * s.cin;
    L = (L + 2 * s.pad - 2 * s.stride) / s.stride + 1;
    mx = mx > (size_t)L * s.cout ? mx : (size_t)L * s.cout;
  }
  return mx;
}

}  // namespace

extern "C" {

#ifdef ACE_PROBE
void ace_debug_set_vae_fused(int on) { g_fused_override = on < 0 ? -1 : (on != 0); }
#endif

size_t ace_vae_packed_bytes(const AceVaeConfig* cfg) {
  if (!cfg || cfg->num_stages < 1 || cfg->num_stages > 8) return 0;
  AceVae tmp;
  tmp.cfg = *cfg;
  return walk_blob(&tmp, nullptr);
}

int ace_vae_create(AceVae** out, const AceVaeConfig* cfg, const void* weights, size_t n_bytes) {
  ACE_REQUIRE(out && cfg && weights, "ace_vae_create: null argument");
  ACE_REQUIRE(cfg->num_stages >= 1 && cfg->num_stages <= 8, "num_stages %d out of range", cfg->num_stages);
  ACE_REQUIRE(cfg->audio_channels == 2, "audio_channels %d unsupported (stereo only)", cfg->audio_channels);
  ACE_REQUIRE(cfg->latent_channels == 64, "latent_channels %d unsupported", cfg->latent_channels);
  ACE_REQUIRE(cfg->encoder_hidden % 128 == 0 && cfg->decoder_channels % 128 == 0,
              "encoder_hidden / decoder_channels must be multiples of 128");
  ACE_REQUIRE(cfg->encoder_hidden == 2 * cfg->latent_channels, "encoder_hidden must be 2*latent_channels");
  for (int i = 0; i < cfg->num_stages; ++i)
    ACE_REQUIRE(cfg->ratios[i] >= 2 && cfg->ratios[i] % 2 == 0 && cfg->ratios[i] <= 2 * 16 - 2,
                "ratio %d unsupported (even strides only)", cfg->ratios[i]);
  int major = 0, dev = 0;
  ACE_CUDA_CHECK(cudaGetDevice(&dev));
  ACE_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  ACE_REQUIRE(major == 10, "libacestep_b200 requires an sm_100 (B200) device, found %d.x", major);
  const size_t need = ace_vae_packed_bytes(cfg);
  ACE_REQUIRE(n_bytes == need, "packed VAE blob has %zu bytes, expected %zu", n_bytes, need);
  AceVae* v = new AceVae();
  v->cfg = *cfg;
  v->hop = 1;
  for (int i = 0; i < cfg->num_stages; ++i) v->hop *= cfg->ratios[i];
  if (cudaMalloc(&v->blob, n_bytes) != cudaSuccess) {
    delete v;
    set_error("cudaMalloc of %zu VAE weight bytes failed", n_bytes);
    return ACE_ERR_NOMEM;
  }
  cudaError_t e = cudaMemcpy(v->blob, weights, n_bytes, cudaMemcpyDefault);
  if (e != cudaSuccess) {
    cudaFree(v->blob);
    delete v;
    set_error("VAE weight upload failed: %s", cudaGetErrorString(e));
    return ACE_ERR_CUDA;
  }
  v->blob_bytes = n_bytes;
  walk_blob(v, v->blob);
  *out = v;
  return ACE_OK;
}

void ace_vae_destroy(AceVae* v) {
  if (!v) return;
  cudaFree(v->blob);
  delete v;
}

// workspace = 5 buffers of (max activation + 2 halos): X, XS (Snake'd X), HS, X2, XS2
size_t ace_vae_decode_workspace_bytes(const AceVae* v, int frames) {
  if (!v || frames <= 0) return 0;
  const size_t per = align256((dec_max_elems(v, frames) + 2 * (size_t)HALO * 2048) * 2);
  return 5 * per + 256;
}
size_t ace_vae_encode_workspace_bytes(const AceVae* v, int samples) {
  if (!v || samples <= 0) return 0;
  const size_t per = align256((enc_max_elems(v, samples) + 2 * (size_t)HALO * 2048) * 2);
  return 5 * per + 256;
}

int ace_vae_decode(AceVae* v, const uint16_t* d_z, int frames, float* d_wav, void* ws, size_t ws_bytes,
                   void* stream) {
  ACE_REQUIRE(v && d_z && d_wav && ws, "ace_vae_decode: null argument");
  ACE_REQUIRE(frames >= 1, "frames %d", frames);
  ACE_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)d_z & 15) == 0, "unaligned buffer");
  ACE_REQUIRE(ws_bytes >= ace_vae_decode_workspace_bytes(v, frames), "decode workspace too small: %zu < %zu",
              ws_bytes, ace_vae_decode_workspace_bytes(v, frames));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t per = align256((dec_max_elems(v, frames) + 2 * (size_t)HALO * 2048) * 2);
  bf16* B[5];
  for (int i = 0; i < 5; ++i) B[i] = reinterpret_cast<bf16*>((uint8_t*)ws + i * per) + (size_t)HALO * 2048;
  bf16 *x = B[0], *xs = B[1], *hs = B[2], *x2 = B[3], *xs2 = B[4];

  long L = frames;
  const int Cz = v->cfg.latent_channels;
  // conv1: latents [L,64] -> [L, C0]; only the Snake'd copy is needed (block 0 starts with Snake -> ConvT)
  {
    const StageW& s0 = v->dec[0];
    int sh[7];
    for (int k = 0; k < 7; ++k) sh[k] = k - 3;
    EpiConv e{nullptr, xs, nullptr, v->dec_conv1.bias, s0.snake.a, s0.snake.ib, (long)s0.cin, 0,
              L * s0.cin, s0.cin};
    ACE_PROPAGATE(conv_gemm((const bf16*)d_z, L, Cz, Cz, v->dec_conv1, s0.cin, L, 7, sh, e, st));
  }
  const int n = v->cfg.num_stages;
  for (int i = 0; i < n; ++i) {
    const StageW& s = v->dec[i];
    // transposed conv: rows q = 0..L (L+1 of them), taps x[q], x[q-1]; output frames q*s - pad + p
    const long Lout = L * s.stride;
    const int sh2[2] = {0, -1};
    EpiConv et{x, xs2, nullptr, s.conv.bias, s.ru[0].s1.a, s.ru[0].s1.ib, (long)s.stride * s.cout,
               -(long)s.pad * s.cout, Lout * s.cout, s.cout};
    ACE_PROPAGATE(conv_gemm(xs, L, s.cin, s.cin, s.conv, s.stride * s.cout, L + 1, 2, sh2, et, st));
    L = Lout;
    // x, xs2 hold the block input; run 3 residual units ping-ponging (x,xs2) <-> (x2,xs)
    const SnakeW& after = (i + 1 < n) ? v->dec[i + 1].snake : v->dec_snake;
    ACE_PROPAGATE(res_unit(s.ru[0], 1, L, s.cout, x, xs2, hs, x2, xs, s.ru[1].s1, st));
    ACE_PROPAGATE(res_unit(s.ru[1], 3, L, s.cout, x2, xs, hs, x, xs2, s.ru[2].s1, st));
    ACE_PROPAGATE(res_unit(s.ru[2], 9, L, s.cout, x, xs2, hs, x2, xs, after, st));
    // block output: x2 (unused further) and xs = Snake_next(x2): the next ConvT's / final conv's input
  }
  const int C = v->cfg.decoder_channels;
  const int smem = 14 * C * 4 + (FC_POS + 6) * C * 2;
  static bool attr = false;
  if (!attr) {
    ACE_CUDA_CHECK(cudaFuncSetAttribute(final_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  prof_begin(PROF_CONV_SIMT, 2.0 * L * 14 * C, (double)L * (C * 2 + 8), st);
  final_conv_kernel<<<(unsigned)((L + FC_POS - 1) / FC_POS), 256, smem, st>>>(xs, v->dec_conv2, d_wav, L, C);
  prof_end(st);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

// Shared body of ace_vae_encode / ace_vae_encode_moments: d_z (posterior sample or mean) and / or d_moments.
static int vae_encode_impl(AceVae* v, const float* d_wav, int samples, const uint16_t* d_eps, uint16_t* d_z,
                           uint16_t* d_moments, void* ws, size_t ws_bytes, void* stream);

int ace_vae_encode(AceVae* v, const float* d_wav, int samples, const uint16_t* d_eps, uint16_t* d_z, void* ws,
                   size_t ws_bytes, void* stream) {
  ACE_REQUIRE(d_z, "ace_vae_encode: null argument");
  return vae_encode_impl(v, d_wav, samples, d_eps, d_z, nullptr, ws, ws_bytes, stream);
}

int ace_vae_encode_moments(AceVae* v, const float* d_wav, int samples, uint16_t* d_moments, void* ws, size_t ws_bytes,
                           void* stream) {
  ACE_REQUIRE(d_moments, "ace_vae_encode_moments: null argument");
  return vae_encode_impl(v, d_wav, samples, nullptr, nullptr, d_moments, ws, ws_bytes, stream);
}

int ace_vae_posterior_sample(const AceVae* v, const uint16_t* d_moments, const uint16_t* d_eps, uint16_t* d_z,
                             int frames, void* stream) {
  ACE_REQUIRE(v && d_moments && d_z && frames >= 1, "ace_vae_posterior_sample: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cz = v->cfg.latent_channels;
  const long tot = (long)frames * Cz;
  prof_begin(PROF_ELEM, 0.0, (double)tot * 8, st);
  posterior_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const bf16*)d_moments, (const bf16*)d_eps, (bf16*)d_z,
                                                                    frames, Cz);
  prof_end(st);
  ACE_CUDA_CHECK(cudaGetLastError());
  return ACE_OK;
}

static int vae_encode_impl(AceVae* v, const float* d_wav, int samples, const uint16_t* d_eps, uint16_t* d_z,
                           uint16_t* d_moments, void* ws, size_t ws_bytes, void* stream) {
  ACE_REQUIRE(v && d_wav && ws, "ace_vae_encode: null argument");
  ACE_REQUIRE(samples >= v->hop && samples % v->hop == 0, "samples %d must be a positive multiple of hop %d",
              samples, v->hop);
  ACE_REQUIRE(((uintptr_t)ws & 255) == 0, "unaligned workspace");
  ACE_REQUIRE(ws_bytes >= ace_vae_encode_workspace_bytes(v, samples), "encode workspace too small: %zu < %zu",
              ws_bytes, ace_vae_encode_workspace_bytes(v, samples));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t per = align256((enc_max_elems(v, samples) + 2 * (size_t)HALO * 2048) * 2);
  bf16* B[5];
  for (int i = 0; i < 5; ++i) B[i] = reinterpret_cast<bf16*>((uint8_t*)ws + i * per) + (size_t)HALO * 2048;
  bf16 *x = B[0], *xs = B[1], *hs = B[2], *x2 = B[3], *xs2 = B[4];
  long L = samples;
  const int H = v->cfg.encoder_hidden;
  const int n = v->cfg.num_stages;
  {
    const SnakeW& s1 = v->enc[0].ru[0].s1;
    ACE_REQUIRE(H % 2 == 0 && 256 % (H / 2) == 0, "encoder_hidden %d unsupported by the first conv kernel", H);
    const long samples_per_block = (long)(256 / (H / 2)) * FC1_S;
    prof_begin(PROF_CONV_SIMT, 2.0 * L * 14 * H, (double)L * (8 + 4.0 * H), st);
    first_conv_kernel<<<(unsigned)((L + samples_per_block - 1) / samples_per_block), 256, 0, st>>>(
        d_wav, v->enc_conv1_w, v->enc_conv1_b, s1.a, s1.ib, x, xs, L, H);
    prof_end(st);
    ACE_CUDA_CHECK(cudaGetLastError());
  }
  for (int i = 0; i < n; ++i) {
    const StageW& s = v->enc[i];
    ACE_PROPAGATE(res_unit(s.ru[0], 1, L, s.cin, x, xs, hs, x2, xs2, s.ru[1].s1, st));
    ACE_PROPAGATE(res_unit(s.ru[1], 3, L, s.cin, x2, xs2, hs, x, xs, s.ru[2].s1, st));
    ACE_PROPAGATE(res_unit(s.ru[2], 9, L, s.cin, x, xs, hs, x2, xs2, s.snake, st));
    // xs2 = snake(block) feeds the strided conv: out[q] = sum_{k<2s} W[:,:,k] xs2[q*s - pad + k].
    // View the zero-haloed buffer, shifted back by `pad` frames, as rows of s frames ([s*Cin]
    // elements): kernel taps 0..s-1 hit view row q, taps s..2s-1 hit view row q+1 -> a 2-tap GEMM.
    const long Lout = (L + 2 * s.pad - 2 * s.stride) / s.stride + 1;
    ACE_CUDA_CHECK(cudaMemsetAsync(xs2 - (size_t)s.pad * s.cin, 0, (size_t)s.pad * s.cin * 2, st));
    ACE_CUDA_CHECK(cudaMemsetAsync(xs2 + (size_t)L * s.cin, 0, (size_t)(s.pad + s.stride) * s.cin * 2, st));
    const SnakeW& next = (i + 1 < n) ? v->enc[i + 1].ru[0].s1 : v->enc_snake;
    EpiConv e{x, xs, nullptr, s.conv.bias, next.a, next.ib, (long)s.cout, 0, Lout * s.cout, s.cout};
    const int sh2[2] = {0, 1};
    ACE_PROPAGATE(conv_gemm(xs2 - (size_t)s.pad * s.cin, Lout + 1, s.stride * s.cin, (long)s.stride * s.cin, s.conv,
                            s.cout, Lout, 2, sh2, e, st));
    L = Lout;
  }
  // conv2 k=3 pad 1 on snake(x): moments [L, 128] -> hs
  {
    const int Cin = v->enc[n - 1].cout;
    const int sh3[3] = {-1, 0, 1};
    ACE_REQUIRE(H == 2 * v->cfg.latent_channels, "encoder_hidden %d != 2 x latent_channels %d", H, v->cfg.latent_channels);
    bf16* mom = d_moments != nullptr ? (bf16*)d_moments : hs;  // [L, 2 Cz]: mean | scale
    EpiConv e{mom, nullptr, nullptr, v->enc_conv2.bias, nullptr, nullptr, (long)H, 0, L * H, H};
    ACE_PROPAGATE(conv_gemm(xs, L, Cin, Cin, v->enc_conv2, H, L, 3, sh3, e, st));
    if (d_z != nullptr) {
      const int Cz = v->cfg.latent_channels;
      const long tot = L * Cz;
      prof_begin(PROF_ELEM, 0.0, (double)tot * 8, st);
      posterior_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(mom, (const bf16*)d_eps, (bf16*)d_z, L, Cz);
      prof_end(st);
      ACE_CUDA_CHECK(cudaGetLastError());
    }
  }
  return ACE_OK;
}

}  // extern "C"
