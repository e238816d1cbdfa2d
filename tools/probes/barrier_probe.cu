// Micro-benchmark of device-wide barrier variants for the persistent decode megakernel (148 co-resident CTAs).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_probe barrier_probe.cu ; run: ./barrier_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// A: arrival = release reduction, wait = acquire poll of the same word
__device__ void bar_a(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    while (ld_acquire(ctr) < epoch) {}
  }
  __syncthreads();
}
// B: returning atomic, last arriver publishes a generation flag on another line
__device__ void bar_b(unsigned* ctr, unsigned* flag, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
    if (old + 1 == gen * gridDim.x) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(gen) : "memory");
    } else {
      while (ld_acquire(flag) < gen) {}
    }
  }
  __syncthreads();
}
// E: every CTA publishes its own generation word; warp 0 of every CTA polls all of them
__device__ void bar_e(unsigned* slots, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x < 32) {
    ++gen;
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(slots + blockIdx.x), "r"(gen) : "memory");
    const unsigned n = gridDim.x;
    bool done = false;
    while (!done) {
      bool ok = true;
      for (unsigned i = threadIdx.x; i < n; i += 32) ok = ok && (ld_relaxed(slots + i) >= gen);
      done = __all_sync(0xffffffffu, ok);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  } else {
    ++gen;
  }
  __syncthreads();
}
// F: like A but the poll is a relaxed load followed by one acquire fence
__device__ void bar_f(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    while (ld_relaxed(ctr) < epoch) {}
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(unsigned* sync, float* data, int iters, long long* cycles) {
  unsigned st = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // a phase worth of traffic: every thread writes one value another CTA reads after the barrier
    data[((size_t)blockIdx.x * 256 + threadIdx.x)] = (float)it;
    if (MODE == 0) bar_a(sync, st);
    else if (MODE == 1) bar_b(sync, sync + 64, st);
    else if (MODE == 2) bar_e(sync + 128, st);
    else if (MODE == 3) bar_f(sync, st);
    else cg::this_grid().sync();
    const float v = data[((size_t)((blockIdx.x + 1) % gridDim.x) * 256 + threadIdx.x)];
    if (v != (float)it) atomicAdd(sync + 32, 1u);     // visibility check
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* sync; float* data; long long* cyc;
  cudaMalloc(&sync, 4096); cudaMalloc(&data, (size_t)sms * 256 * 4); cudaMalloc(&cyc, 8);
  const char* names[] = {"A red.release + acquire poll (same word)", "B atom.acq_rel + flag line", "E per-CTA slots, warp poll",
                         "F red.release + relaxed poll + fence", "G cooperative_groups grid.sync"};
  void* fns[] = {(void*)probe<0>, (void*)probe<1>, (void*)probe<2>, (void*)probe<3>, (void*)probe<4>};
  int iters = 2000;
  for (int m = 0; m < 5; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(sync, 0, 4096);
      void* args[] = {&sync, &data, &iters, &cyc};
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      cudaError_t e = cudaLaunchCooperativeKernel(fns[m], dim3(sms), dim3(256), args, 0, 0);
      cudaEventRecord(b);
      cudaError_t e2 = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, a, b);
      unsigned bad = 0; cudaMemcpy(&bad, sync + 32, 4, cudaMemcpyDeviceToHost);
      if (rep == 1)
        printf("%-45s %7.3f us/iter  (stale reads %u, launch %d sync %d)\n", names[m], ms * 1000.0 / iters, bad, (int)e, (int)e2);
    }
  }
  return 0;
}
