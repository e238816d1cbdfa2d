// Streaming probe for the decode-attention ring: how fast can W warps per SM each pull a private contiguous slab
// from HBM through `nslot` bulk copies (cp.async.bulk, mbarrier completion) of `unit` bytes, with and without the
// consumer arithmetic of the score pass?  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ring_probe ring_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: wait + refill only (one 16-byte read per lane so the data is touched); 1: score-pass arithmetic (8 lanes per
// 128-byte row, butterfly); 2: PV-style arithmetic (8 accumulators per lane, no shuffles)
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(const uint8_t* __restrict__ data, size_t slab_bytes, int warps, int nslot,
                                                int unit, float* sink, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(bars + i, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const long long t0 = clock64();
  float accs = 0.f;
  if (warp < warps) {
    const uint8_t* src = data + ((size_t)blockIdx.x * warps + warp) * slab_bytes;
    uint8_t* ring = smem + (size_t)warp * nslot * unit;
    uint64_t* wb = bars + warp * 8;
    const int units = (int)(slab_bytes / unit);
    if (lane == 0) for (int u = 0; u < nslot && u < units; ++u) { mbar_expect_tx(wb + u, unit); bulk_g2s(ring + (size_t)u * unit, src + (size_t)u * unit, unit, wb + u); }
    uint32_t par = 0; int slot = 0;
    const int sub = lane & 7, rslot = lane >> 3;
    float q[8]; for (int i = 0; i < 8; ++i) q[i] = 0.01f * (sub + i);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float mx = -1e30f;
    for (int u = 0; u < units; ++u) {
      while (!mbar_try_wait(wb + slot, (par >> slot) & 1u)) {}
      par ^= 1u << slot;
      const uint8_t* s = ring + (size_t)slot * unit;
      for (int c = 0; c < unit; c += 4096) {
        uint4 kv[8];
        if (MODE == 0) { kv[0] = *reinterpret_cast<const uint4*>(s + c + lane * 16); accs += __uint_as_float(kv[0].x & 0x3f800000u); }
        else {
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) kv[ps] = *reinterpret_cast<const uint4*>(s + c + rslot * 128 + sub * 16 + ps * 512);
        }
        if (c + 4096 >= unit) {
          __syncwarp();
          if (lane == 0 && u + nslot < units) { mbar_expect_tx(wb + slot, unit); bulk_g2s(ring + (size_t)slot * unit, src + (size_t)(u + nslot) * unit, unit, wb + slot); }
        }
        if (MODE == 1) {
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[ps]);
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h2[i]); part = fmaf(q[2 * i], f.x, part); part = fmaf(q[2 * i + 1], f.y, part); }
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            mx = fmaxf(mx, part);
          }
        } else if (MODE == 2) {
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[ps]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h2[i]); acc[2 * i] = fmaf(q[ps], f.x, acc[2 * i]); acc[2 * i + 1] = fmaf(q[ps], f.y, acc[2 * i + 1]); }
          }
        }
      }
      slot = slot + 1 == nslot ? 0 : slot + 1;
    }
    for (int i = 0; i < 8; ++i) accs += acc[i];
    accs += mx;
  }
  if (accs == 123.456f) sink[0] = accs;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t slab = 640 * 1024;                 // four attention items (~630 positions x 256 B each) per warp
  const size_t total = (size_t)sms * 8 * slab * 12;   // 12 "layers" worth so that nothing stays in L2 (1.4 GB+)
  uint8_t* data; cudaMalloc(&data, total); cudaMemset(data, 0x3c, total);
  float* sink; cudaMalloc(&sink, 4); long long* cyc; cudaMalloc(&cyc, 8 * 256);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int warps, nslot, unit; };
  Cfg cfgs[] = {{6, 5, 4096}, {6, 8, 4096}, {6, 2, 4096}, {6, 4, 8192}, {6, 2, 16384}, {8, 6, 4096}, {8, 3, 8192}, {4, 8, 4096}, {4, 6, 8192},
                {2, 8, 8192}, {1, 8, 16384}, {5, 6, 4096}};
  printf("HBM streaming through bulk-copy rings, %d SMs, slab %zu KB per warp\n", sms, slab / 1024);
  for (int mode = 0; mode < 3; ++mode) {
    for (auto c : cfgs) {
      if ((size_t)c.warps * c.nslot * c.unit > 196 * 1024) continue;
      float best = 1e9f;
      for (int rep = 0; rep < 12; ++rep) {      // each rep reads a different 1/12 of the buffer
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const uint8_t* base = data + (size_t)rep * sms * 8 * slab;
        cudaEventRecord(a);
        if (mode == 0) probe<0><<<sms, 256, 200 * 1024>>>(base, slab, c.warps, c.nslot, c.unit, sink, cyc);
        else if (mode == 1) probe<1><<<sms, 256, 200 * 1024>>>(base, slab, c.warps, c.nslot, c.unit, sink, cyc);
        else probe<2><<<sms, 256, 200 * 1024>>>(base, slab, c.warps, c.nslot, c.unit, sink, cyc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2 && ms < best) best = ms;
        cudaEventDestroy(a); cudaEventDestroy(b);
      }
      cudaError_t e = cudaGetLastError();
      const double bytes = (double)sms * c.warps * slab;
      printf("mode %d  warps %d  nslot %d  unit %5d : %7.2f us  %6.2f TB/s  (%s)\n", mode, c.warps, c.nslot, c.unit, best * 1e3,
             bytes / (best * 1e-3) / 1e12, cudaGetErrorString(e));
    }
  }
  return 0;
}
