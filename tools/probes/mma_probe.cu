// Micro-benchmarks behind two design decisions (built and run on the GPU box; output committed under profiles/):
//   1. cost of ONE tcgen05.mma (kind::f16, K = 16) as a function of its shape and operand source -- the decode
//      megakernel's GEMM phases are bound by the NUMBER of MMA instructions (profiles/r01/mega_phase_breakdown_v14);
//   2. peak fp32 FMA rate of the chip (the denominator of the VQ-argmin kernel, SURVEY 8d: "run an FFMA micro-benchmark").
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../ivideogpt_b200/csrc -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace ivg {      // common.cuh declares these; the probe is stand-alone
void set_error(const char*, ...) {}
const char* last_error() { return ""; }
unsigned long long g_launches = 0;
bool g_pdl = false;
}
using namespace ivg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// A operand from TMEM (tcgen05.mma "TS" form)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: SS, one accumulator; 1: SS, two accumulators alternating; 2: TS (A in TMEM), one accumulator;
// mode 3 / 4: SS with 2 / 4 ISSUING WARPS (lane 0 of warps 0..W-1), each with its own accumulator (N <= 128) and its own share
// of the nmma instructions -- does the ~112-cycle floor belong to the issuing thread or to the tensor pipe?
// mode 5 / 6: 2 / 4 issuing warps accumulating into the SAME accumulator (operands are all ones, so every element of D must end
// at exactly 16 * nmma): is concurrent accumulation from several threads safe, and how fast is it?
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int M, int N, int nmma, int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a = base;                       // 4 k-blocks x 128 rows x 128 B = 64 KB
  uint8_t* b = base + 64 * 1024;           // 4 k-blocks x 256 rows x 128 B = 128 KB
  __shared__ uint64_t bar[4];
  __shared__ uint32_t holder;
  __shared__ int s_bad;
  const uint32_t fill = mode >= 5 ? 0x3F803F80u : 0u;      // bf16 1.0 pairs
  for (int i = threadIdx.x; i < (192 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(fill, fill, fill, fill);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1); fence_barrier_init(); s_bad = 0; }
  if (threadIdx.x < 32) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  const int issuers = (mode == 3 || mode == 5) ? 2 : ((mode == 4 || mode == 6) ? 4 : 1);
  const int w = threadIdx.x >> 5;
  if (mode >= 5) {
    // zero the accumulator first with one MMA that does not accumulate, issued by warp 0, then let every issuer accumulate
    if (threadIdx.x == 0) {
      const uint32_t idesc = umma_idesc(1, M, N);
      umma_ss<false>(tm, umma_desc_sw128_kmajor(smem_u32(a)), umma_desc_sw128_kmajor(smem_u32(b)), idesc, 0u);
      umma_commit(bar + 3);
      mbar_wait(bar + 3, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    const uint32_t idesc = umma_idesc(1, M, N);
    const uint32_t a0 = smem_u32(a), b0 = smem_u32(b);
    const int mine = nmma / issuers;
    const long long t0 = clock64();
    for (int i = 0; i < mine; ++i) {
      const int kb = (i >> 2) & 3, ks = i & 3;
      const uint64_t ad = umma_desc_sw128_kmajor(a0 + kb * (128 * 128)) + (uint64_t)(ks * 2);
      const uint64_t bd = umma_desc_sw128_kmajor(b0 + kb * (256 * 128)) + (uint64_t)(ks * 2);
      if (mode == 2) umma_ts(tm, tm + 256 + (uint32_t)((i & 15) * 8), bd, idesc, i > 0);
      else if (mode >= 5) umma_ss<false>(tm, ad, bd, idesc, 1u);
      else if (mode >= 3) umma_ss<false>(tm + (uint32_t)(w * 128), ad, bd, idesc, i > 0 ? 1u : 0u);
      else umma_ss<false>(tm + (mode == 1 ? (uint32_t)((i & 1) * 256) : 0u), ad, bd, idesc, i > 1 ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit(bar + w);
    mbar_wait(bar + w, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && w == issuers - 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (mode >= 5) {       // every element must be 16 * (nmma + 1): the zeroing MMA contributed one K = 16 step too
    const float want = 16.0f * (float)(nmma + 1);
    int bad = 0;
    const int q = threadIdx.x >> 5;
    const bool lane_has_row = M == 128 || (threadIdx.x & 31) < 16;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      if (lane_has_row)
        for (int i = 0; i < 16; ++i) bad += (__uint_as_float(r[i]) != want);
    }
    if (bad) atomicAdd(&s_bad, bad);
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[2] = s_bad;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// ---- fp32 FMA peak: 16 independent chains per thread, register-register-register and register-immediate forms ----
template <int IMM>
__global__ void __launch_bounds__(1024) ffma_probe_kernel(float* sink, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = (float)(threadIdx.x + j) * 1e-3f;
  a += (float)threadIdx.x * 1e-10f;      // thread-dependent: keeps the multiplier in a vector register (no uniform-register form)
  b += (float)threadIdx.x * 1e-12f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (IMM) asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, %1;" : "+f"(x[j]) : "f"(b));
      else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(a), "f"(b));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += x[j];
  if (s == 123.456f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  printf("{\"sms\": %d, \"clock_khz_nominal\": %d,\n", sms, khz);
  long long* out;
  CK(cudaMalloc(&out, 64));
  CK(cudaMemset(out, 0, 64));
  const size_t smem = 193 * 1024 + 1024;
  CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf(" \"mma\": [\n");
  const int Ms[2] = {64, 128};
  const int Ns[7] = {8, 16, 32, 64, 128, 256, 48};
  bool first = true;
  for (int mode = 0; mode < 7; ++mode)
    for (int mi = 0; mi < 2; ++mi)
      for (int ni = 0; ni < 7; ++ni) {
        const int M = Ms[mi], N = Ns[ni], nm = 512;
        if (mode == 2 && M == 64) continue;              // TS form: M = 128 only here
        if ((mode == 3 || mode == 4) && N > 128) continue;
        if (mode >= 5 && N < 64) continue;
        long long h[3] = {0, 0, 0};
        for (int rep = 0; rep < 3; ++rep) {              // last repetition is reported (warm)
          mma_probe_kernel<<<sms, 128, smem>>>(M, N, nm, mode, out);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost));
        }
        const char* names[7] = {"SS", "SS-2acc", "TS", "SS-2warps", "SS-4warps", "SS-2warps-same-acc", "SS-4warps-same-acc"};
        printf("%s  {\"mode\": \"%s\", \"M\": %d, \"N\": %d, \"issue_cycles_per_mma\": %.1f, \"complete_cycles_per_mma\": %.1f, \"wrong_elements\": %lld}",
               first ? "" : ",\n", names[mode], M, N, (double)h[0] / nm, (double)h[1] / nm, mode >= 5 ? h[2] : 0LL);
        first = false;
      }
  printf("\n ],\n");
  float* sink;
  CK(cudaMalloc(&sink, (size_t)sms * 8 * 1024 * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 20000;
  printf(" \"ffma\": {");
  for (int imm = 0; imm < 2; ++imm) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      if (imm) ffma_probe_kernel<1><<<sms * 2, 1024>>>(sink, iters, 0.999f, 1e-7f);
      else ffma_probe_kernel<0><<<sms * 2, 1024>>>(sink, iters, 0.999f, 1e-7f);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    const double flops = 2.0 * 16.0 * iters * (double)sms * 2 * 1024;
    printf("%s\"%s_tflops\": %.2f, \"%s_ms\": %.3f", imm ? ", " : "", imm ? "reg_imm" : "reg_reg", flops / (best * 1e-3) / 1e12,
           imm ? "reg_imm" : "reg_reg", best);
  }
  printf("}\n}\n");
  return 0;
}
