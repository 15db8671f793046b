// gemm_probe.cu — stand-alone bring-up probe for the tcgen05 GEMM (no torch, starts in <1 s).
// Checks the tensor-core path and the scalar debug path against a host fp32 loop on a few
// shapes (plain linear, ragged M, 7-tap dilated conv, 2-tap transposed-conv form) and prints
// timing for the DiT-sized problems.  Build: see tools/build_probe.sh.  Dev tool, not shipped.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../ace-step-1.5-for-windows_b200/csrc/epilogues.cuh"
#include "../ace-step-1.5-for-windows_b200/csrc/gemm.cuh"

using namespace ace;

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t r = u + 0x7fff + ((u >> 16) & 1);
  return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static int run_case(const char* name, int M, int N, int kc, int ntaps, const int* shifts,
                    int a_rows, bool ref_path, int check_stride, int timing_iters, int bn = 128, bool splitk = false) {
  const long Ktot = (long)ntaps * kc;
  std::vector<uint16_t> hA((size_t)a_rows * kc), hB((size_t)N * Ktot), hbias(N);
  for (auto& x : hA) x = f2bf(frand());
  for (auto& x : hB) x = f2bf(frand() * 0.05f);
  for (auto& x : hbias) x = f2bf(frand());
  bf16 *dA, *dB, *dbias, *dO;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dbias, hbias.size() * 2));
  CK(cudaMalloc(&dO, (size_t)M * N * 2));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), hbias.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, (size_t)M * N * 2));
  GemmPlan plan;
  if (make_gemm_plan(&plan, dA, a_rows, kc, kc, dB, N, Ktot, M, ntaps, shifts, bn) != ACE_OK) {
    printf("[%s] plan failed: %s\n", name, get_error());
    return 1;
  }
  void* scratch = nullptr;
  if (splitk) {  // let the cost model decide; report what it chose
    CK(cudaMalloc(&scratch, gemm_splitk_scratch_bytes()));
    CK(cudaMemset(scratch, 0, gemm_splitk_scratch_bytes()));
    if (gemm_plan_enable_splitk(&plan, scratch, gemm_splitk_scratch_bytes()) != ACE_OK) {
      printf("[%s] split-K plan failed: %s\n", name, get_error());
      return 1;
    }
    printf("[%s] split-K: bn=%d splits=%d\n", name, plan.bn, plan.shp.splits);
    bn = plan.bn;
  }
  EpiBias epi{dO, (long)N, dbias};
  set_gemm_debug_reference(ref_path);
  if (launch_gemm(plan, epi, 0) != ACE_OK) {
    printf("[%s] launch failed: %s\n", name, get_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] kernel failed: %s\n", name, cudaGetErrorString(e));
    return 1;
  }
  std::vector<uint16_t> hO((size_t)M * N);
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  long checked = 0;
  for (long idx = 0; idx < (long)M * N; idx += check_stride) {
    int m = idx / N, n = idx % N;
    double acc = bf2f(hbias[n]);
    for (int t = 0; t < ntaps; ++t) {
      long r = m + (shifts ? shifts[t] : 0);
      if (r < 0 || r >= a_rows) continue;
      for (int k = 0; k < kc; ++k)
        acc += (double)bf2f(hA[r * kc + k]) * bf2f(hB[(long)n * Ktot + (long)t * kc + k]);
    }
    double got = bf2f(hO[idx]);
    double err = fabs(got - acc);
    if (err > max_err) max_err = err;
    if (fabs(acc) > max_ref) max_ref = fabs(acc);
    ++checked;
  }
  const bool pass = max_err <= 0.02 * (max_ref > 1 ? max_ref : 1);
  printf("[%s] %s bn=%d M=%d N=%d Kc=%d taps=%d : checked=%ld max_err=%.4g max_ref=%.4g -> %s\n", name,
         ref_path ? "ref" : "tc ", bn, M, N, kc, ntaps, checked, max_err, max_ref, pass ? "PASS" : "FAIL");
  if (pass && timing_iters > 0 && !ref_path) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch_gemm(plan, epi, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < timing_iters; ++i) launch_gemm(plan, epi, 0);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= timing_iters;
    double fl = 2.0 * M * N * (double)Ktot;
    printf("[%s]    %.3f us/launch  %.1f TFLOP/s\n", name, ms * 1e3, fl / ms * 1e-9);
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dbias);
  cudaFree(dO);
  if (scratch) cudaFree(scratch);
  return pass ? 0 : 1;
}

int main() {
  int fails = 0;
  const int one[1] = {0};
  // scalar debug path first (validates plans/epilogue independent of tcgen05)
  fails += run_case("lin-small", 300, 256, 128, 1, one, 300, true, 7, 0);
  fails += run_case("lin-small", 300, 256, 128, 1, one, 300, false, 1, 0);
  fails += run_case("lin-1tile", 128, 128, 64, 1, one, 128, false, 1, 0);
  fails += run_case("lin-k2048", 1500, 2048, 2048, 1, one, 1500, false, 97, 20);
  fails += run_case("lin-qkv", 1500, 4096, 2048, 1, one, 1500, false, 997, 20);
  fails += run_case("lin-gateup", 1500, 12288, 2048, 1, one, 1500, false, 4999, 20);
  fails += run_case("lin-down", 1500, 2048, 6144, 1, one, 1500, false, 997, 20);
  const int conv7[7] = {-27, -18, -9, 0, 9, 18, 27};
  fails += run_case("conv7-d9", 1000, 128, 128, 7, conv7, 1000, true, 13, 0);
  fails += run_case("conv7-d9", 1000, 128, 128, 7, conv7, 1000, false, 3, 0);
  const int convt[2] = {0, -1};
  fails += run_case("convT", 501, 1024, 256, 2, convt, 500, false, 11, 0);
  fails += run_case("conv7-big", 96000, 128, 128, 7, conv7, 96000, false, 9973, 10);
  // CTA-pair 256 x 256 tiles
  fails += run_case("pair-1tile", 256, 256, 64, 1, one, 256, false, 1, 0, 256);
  fails += run_case("pair-small", 300, 384, 128, 1, one, 300, false, 1, 0, 256);
  fails += run_case("pair-k2048", 1500, 2048, 2048, 1, one, 1500, false, 97, 20, 256);
  fails += run_case("pair-qkv", 1500, 4096, 2048, 1, one, 1500, false, 997, 20, 256);
  fails += run_case("pair-gateup", 1500, 12288, 2048, 1, one, 1500, false, 4999, 20, 256);
  fails += run_case("pair-down", 1500, 2048, 6144, 1, one, 1500, false, 997, 20, 256);
  fails += run_case("pair-conv7", 5000, 512, 512, 7, conv7, 5000, false, 1013, 10, 256);
  fails += run_case("pair-conv7-n128", 96000, 128, 128, 7, conv7, 96000, false, 9973, 10, 256);
  // split-K (few output tiles): every element checked on the small ones, launched repeatedly (counter re-arm)
  fails += run_case("splitk-m125", 125, 2048, 2048, 1, one, 125, false, 1, 20, 0, true);
  fails += run_case("nosplit-m125", 125, 2048, 2048, 1, one, 125, false, 97, 20, 0, false);
  fails += run_case("splitk-qkv-m125", 125, 4096, 2048, 1, one, 125, false, 3, 20, 0, true);
  fails += run_case("splitk-gateup-m125", 125, 12288, 2048, 1, one, 125, false, 997, 20, 0, true);
  fails += run_case("splitk-down-m125", 125, 2048, 6144, 1, one, 125, false, 3, 20, 0, true);
  fails += run_case("splitk-m375", 375, 2048, 2048, 1, one, 375, false, 7, 20, 0, true);
  fails += run_case("nosplit-m375", 375, 2048, 2048, 1, one, 375, false, 97, 20, 0, false);
  fails += run_case("splitk-m40-n128", 40, 128, 2048, 1, one, 40, false, 1, 20, 0, true);
  fails += run_case("splitk-conv7", 200, 256, 128, 7, conv7, 200, false, 1, 0, 0, true);
  printf("gemm_probe: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}
