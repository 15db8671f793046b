// runtime.cu — host-side plumbing shared by every entry point of libacestep_b200:
// error strings, TMA tensor-map encoding (driver entry point resolved at run time so the
// library has no link-time dependency on libcuda), GEMM plans, and the scalar reference GEMM
// used only by the bring-up/debug switch.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "gemm.cuh"

namespace ace {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- launch accounting + per-launch event profiling -------------------------------------------
struct ProfRec {
  int cat;
  cudaEvent_t a, b;
  double flops, bytes;
  int m, n, k;  // GEMM launches only (0 otherwise)
};
struct ShapeAgg {
  int m, n, k, launches;
  double ms;
};
static std::vector<ShapeAgg> g_shapes;  // per-(M,N,K) totals of the last profiling window
static int g_tag_m = 0, g_tag_n = 0, g_tag_k = 0;
void prof_tag_gemm(int m, int n, int k) { g_tag_m = m; g_tag_n = n; g_tag_k = k; }
static uint64_t g_launches = 0;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

void add_launches(uint64_t n) { g_launches += n; }
uint64_t launch_count() { return g_launches; }
bool prof_active() { return g_prof_on; }
void prof_begin(int cat, double flops, double bytes, cudaStream_t st) {
  ++g_launches;
  if (!g_prof_on) return;
  ProfRec r{cat, nullptr, nullptr, flops, bytes, 0, 0, 0};
  if (cat == PROF_GEMM) {
    r.m = g_tag_m;
    r.n = g_tag_n;
    r.k = g_tag_k;
  }
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().b, st);
}
void prof_start() {
  for (ProfRec& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on = true;
}
// Sums per category: ms, flops, bytes, launches (arrays of PROF_NCAT).  Synchronises the device.
int prof_stop(float* ms, double* flops, double* bytes, int* launches) {
  g_prof_on = false;
  cudaError_t e = cudaDeviceSynchronize();
  for (int c = 0; c < PROF_NCAT; ++c) {
    ms[c] = 0.f;
    flops[c] = bytes[c] = 0.0;
    launches[c] = 0;
  }
  g_shapes.clear();
  for (ProfRec& r : g_prof) {
    float t = 0.f;
    if (e == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cat >= 0 &&
        r.cat < PROF_NCAT) {
      ms[r.cat] += t;
      flops[r.cat] += r.flops;
      bytes[r.cat] += r.bytes;
      launches[r.cat] += 1;
      if (r.cat == PROF_GEMM) {
        bool found = false;
        for (ShapeAgg& s : g_shapes)
          if (s.m == r.m && s.n == r.n && s.k == r.k) {
            s.launches += 1;
            s.ms += t;
            found = true;
            break;
          }
        if (!found) g_shapes.push_back(ShapeAgg{r.m, r.n, r.k, 1, (double)t});
      }
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  if (e != cudaSuccess) {
    set_error("profile stop: %s", cudaGetErrorString(e));
    return ACE_ERR_CUDA;
  }
  return ACE_OK;
}

// Top GEMM shapes (by total time) of the last profiling window; returns how many were written.
int prof_gemm_shapes(int max_out, int* m, int* n, int* k, int* launches, float* ms) {
  std::vector<ShapeAgg> v = g_shapes;
  for (size_t i = 0; i < v.size(); ++i)
    for (size_t j = i + 1; j < v.size(); ++j)
      if (v[j].ms > v[i].ms) std::swap(v[i], v[j]);
  int cnt = 0;
  for (const ShapeAgg& s : v) {
    if (cnt >= max_out) break;
    m[cnt] = s.m;
    n[cnt] = s.n;
    k[cnt] = s.k;
    launches[cnt] = s.launches;
    ms[cnt] = (float)s.ms;
    ++cnt;
  }
  return cnt;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// L2 promotion of the tensor maps (experiments: ACE_TMAP_L2=0|64|128|256, default 256)
static CUtensorMapL2promotion tmap_l2_promotion() {
  static int v = -1;
  if (v < 0) {
    const char* e = probe_env("ACE_TMAP_L2");
    v = e ? atoi(e) : 256;
  }
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                          : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                   uint64_t row_pitch_bytes, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return ACE_ERR_CUDA;
  }
  ACE_REQUIRE(((uintptr_t)base & 15) == 0, "tensor map base %p not 16-byte aligned", base);
  ACE_REQUIRE((row_pitch_bytes & 15) == 0, "tensor map pitch %llu not a multiple of 16",
              (unsigned long long)row_pitch_bytes);
  ACE_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map box rows %u", box_rows);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {GEMM_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  tmap_l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p cols=%llu rows=%llu pitch=%llu box=%u",
              (int)r, base, (unsigned long long)cols, (unsigned long long)rows,
              (unsigned long long)row_pitch_bytes, box_rows);
    return ACE_ERR_CUDA;
  }
  return ACE_OK;
}

int encode_tmap_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batches,
                   uint64_t row_pitch_bytes, uint64_t batch_pitch_bytes, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return ACE_ERR_CUDA;
  }
  ACE_REQUIRE(((uintptr_t)base & 15) == 0, "tensor map base %p not 16-byte aligned", base);
  ACE_REQUIRE((row_pitch_bytes & 15) == 0 && (batch_pitch_bytes & 15) == 0, "tensor map pitch not a multiple of 16");
  cuuint64_t dims[3] = {cols, rows, batches};
  cuuint64_t strides[2] = {row_pitch_bytes, batch_pitch_bytes};
  cuuint32_t box[3] = {GEMM_BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult res = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d): cols=%llu rows=%llu batches=%llu", (int)res,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batches);
    return ACE_ERR_CUDA;
  }
  return ACE_OK;
}

int make_gemm_plan(GemmPlan* plan, const bf16* a, int a_rows, int kc, long a_ld, const bf16* b,
                   int n, long b_ld, int m, int ntaps, const int* shifts, int bn) {
  ACE_REQUIRE(kc > 0 && kc % GEMM_BK == 0, "gemm: per-tap K %d must be a multiple of %d", kc,
              GEMM_BK);
  ACE_REQUIRE(ntaps >= 1 && ntaps <= GEMM_MAX_TAPS, "gemm: ntaps %d", ntaps);
  if (bn == 0) {
    // auto: CTA-pair 256 x 256 tiles once the problem spans more than one 128-wide tile in both
    // directions (halves L2 operand traffic per FLOP); ACE_GEMM_BN=128|256 overrides for experiments
    static int forced = -1;
    if (forced < 0) {
      const char* e = probe_env("ACE_GEMM_BN");
      forced = e ? atoi(e) : 0;
    }
    bn = forced ? forced : ((n >= 256 && m > 128) ? 256 : 128);
  } else if (bn == -192) {
    // auto among {128, 256, 192}: the caller's epilogue accepts 64-column sub-tiles.  Cost model:
    // waves of CTA pairs x tile width (the main loop time of one pair tile is proportional to bn).
    static int forced = -1;
    if (forced < 0) {
      const char* e = probe_env("ACE_GEMM_BN");
      forced = e ? atoi(e) : 0;
    }
    if (forced) {
      bn = forced;
    } else if (!(n >= 256 && m > 128)) {
      bn = 128;
    } else {
      const int clusters = num_sms() / 2, mt = ceil_div(m, 256);
      const long c256 = (long)ceil_div(mt * ceil_div(n, 256), clusters) * 256;
      const long c192 = (long)ceil_div(mt * ceil_div(n, 192), clusters) * 192;
      bn = c192 < c256 ? 192 : 256;
    }
  }
  ACE_REQUIRE(bn == 128 || bn == 256 || bn == 192, "gemm: BLOCK_N %d unsupported", bn);
  memset(plan, 0, sizeof(*plan));
  plan->shp.M = m;
  plan->shp.N = n;
  plan->shp.kblocks_per_tap = kc / GEMM_BK;
  plan->shp.ntaps = ntaps;
  plan->shp.b_tap_stride = kc;
  for (int i = 0; i < ntaps; ++i) plan->shp.shift[i] = shifts ? shifts[i] : 0;
  plan->a_ptr = a;
  plan->a_ld = a_ld;
  plan->a_rows = a_rows;
  plan->b_ptr = b;
  plan->b_ld = b_ld;
  plan->bn = bn;
  {
    // stream the larger operand once, keep the smaller one in L2 (see GemmShape::raster_n)
    static int forced = -1;
    if (forced < 0) {
      const char* e = probe_env("ACE_GEMM_RASTER");
      forced = e ? atoi(e) : 2;
    }
    // measured rule (tools/time_step.py A/B): only where the m-fastest order demonstrably re-reads A from HBM once
    // per wave — a plain GEMM whose A exceeds what L2 keeps next to B (down_proj at M = 6000: 74 MB, ncu 308 MB of
    // DRAM traffic for 172 MB of operands); the codec's tap-shifted convolutions stay m-fastest
    const double a_bytes = 2.0 * m * kc, b_bytes = 2.0 * n * kc * ntaps;
    plan->shp.raster_n = forced != 2 ? forced : (ntaps == 1 && a_bytes > 48e6 && a_bytes > b_bytes ? 1 : 0);
  }
  {
    plan->shp.race_delay = probe_env("ACE_RACE_DELAY") != nullptr;  // read per plan: a test flips it per handle
  }
  if (m <= 0 || n <= 0) return ACE_OK;
  ACE_PROPAGATE(encode_tmap_2d(&plan->tma_a, a, (uint64_t)kc, (uint64_t)a_rows,
                               (uint64_t)a_ld * sizeof(bf16), GEMM_BM));
  // B is staged in 128-row boxes (the 256-wide pair kernel loads one half of N per CTA), 96-row
  // boxes for the 192-wide pair tile
  ACE_PROPAGATE(encode_tmap_2d(&plan->tma_b, b, (uint64_t)ntaps * kc, (uint64_t)n,
                               (uint64_t)b_ld * sizeof(bf16), bn == 192 ? 96u : 128u));
  return ACE_OK;
}

int gemm_plan_enable_splitk(GemmPlan* plan, void* scratch, size_t scratch_bytes) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = probe_env("ACE_NO_SPLITK");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  const GemmShape& sh = plan->shp;
  if (disabled || scratch == nullptr || scratch_bytes < gemm_splitk_scratch_bytes() || sh.M <= 0 || sh.N <= 0 ||
      sh.N % 128 != 0)
    return ACE_OK;
  const int total_kb = sh.ntaps * sh.kblocks_per_tap, sms = num_sms();
  const int tiles = ceil_div(sh.M, GEMM_BM) * (sh.N / 128);
  int best = 1;
  for (int s = 2; s <= 8; s *= 2)
    if (tiles * s <= sms && tiles * s <= GEMM_SPLITK_MAX_WORKS && total_kb / s >= 2) best = s;
  if (best == 1) return ACE_OK;
  // Cost = operand KB one CTA has to ingest on the critical path (the main loops are TMA-ingest bound,
  // ~57 B/clk/SM); the split-K tail (partial store, inter-CTA wait, reduce) is charged as 24 K blocks
  // (~7 us): measured on B200, a 2-way split of M = 375 x N = 2048 x K = 2048 loses (16.8 vs 15.3 us)
  // while 8-way splits at M = 125 win once the weights come from HBM rather than L2.
  const long cost_split = (long)ceil_div(total_kb, best) * 32 + 24 * 32;
  long cost_now;
  if (plan->bn == 128) {
    cost_now = (long)ceil_div(tiles, sms) * total_kb * 32;
  } else {
    const int pair_tiles = ceil_div(sh.M, 256) * ceil_div(sh.N, plan->bn);
    cost_now = (long)ceil_div(pair_tiles, sms / 2) * total_kb * (plan->bn == 192 ? 28 : 32);
  }
  if (cost_split >= cost_now) return ACE_OK;
  if (plan->bn == 192)  // the single-CTA kernel stages B in 128-row boxes
    ACE_PROPAGATE(encode_tmap_2d(&plan->tma_b, plan->b_ptr, (uint64_t)sh.ntaps * sh.b_tap_stride, (uint64_t)sh.N,
                                 (uint64_t)plan->b_ld * sizeof(bf16), 128u));
  plan->bn = 128;
  plan->shp.splits = best;
  plan->shp.part = reinterpret_cast<float*>(scratch);
  plan->shp.sync = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(scratch) +
                                               (size_t)GEMM_SPLITK_MAX_WORKS * GEMM_BM * 128 * sizeof(float));
  return ACE_OK;
}

#ifdef ACE_PROBE
static bool g_debug_ref = false;
void set_gemm_debug_reference(bool on) { g_debug_ref = on; }
bool gemm_debug_reference() { return g_debug_ref; }

float* gemm_debug_scratch(size_t elems) {
  static float* buf = nullptr;
  static size_t cap = 0;
  if (elems > cap) {
    if (buf) {
      cudaDeviceSynchronize();
      cudaFree(buf);
      buf = nullptr;
      cap = 0;
    }
    if (cudaMalloc(&buf, elems * sizeof(float)) != cudaSuccess) {
      buf = nullptr;
      return nullptr;
    }
    cap = elems;
  }
  return buf;
}

__global__ void gemm_ref_kernel(const bf16* __restrict__ A, long lda, int a_rows,
                                const bf16* __restrict__ B, long ldb, GemmShape shp, int kc,
                                float* __restrict__ out, int ld_out, int m_pad, int n_pad) {
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int m = blockIdx.y * 8 + threadIdx.y;
  if (m >= m_pad || n >= n_pad) return;
  float acc = 0.f;
  if (m < shp.M && n < shp.N) {
    for (int tap = 0; tap < shp.ntaps; ++tap) {
      const long r = (long)m + shp.shift[tap];
      if (r < 0 || r >= a_rows) continue;
      const bf16* ap = A + r * lda;
      const bf16* bp = B + (long)n * ldb + (long)tap * shp.b_tap_stride;
      for (int k = 0; k < kc; ++k) acc += __bfloat162float(ap[k]) * __bfloat162float(bp[k]);
    }
  }
  out[(size_t)m * ld_out + n] = acc;
}

#endif  // ACE_PROBE

}  // namespace ace

#ifdef ACE_PROBE
namespace ace {
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ACE_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
}  // namespace ace
#endif
