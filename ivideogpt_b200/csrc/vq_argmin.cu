// Fused L2-distance argmin over a VQ codebook (replaces torch.cdist + argmin inside the diffusers
// VectorQuantizer called at reference ivideogpt/vq_model/compressive_vq_model.py:199,202).
//
// Definition computed here (and bit-for-bit by oracle/vq_argmin_ref.c):
//     dot(n,k)   = fma-chain over d = 0..D-1 in order, starting from +0.0f:  acc = fmaf(z[n][d], e[k][d], acc)
//     enorm(k)   = fma-chain over d in order:                                   s   = fmaf(e[k][d], e[k][d], s)
//     score(n,k) = fmaf(-2.0f, dot(n,k), enorm(k))          ( == ||z-e||^2 - ||z||^2 )
//     idx(n)     = the smallest k attaining min_k score(n,k)   (torch.argmin tie rule)
// ||z||^2, the clamp and the sqrt of torch.cdist are monotone per row and do not change the argmin
// except on exact/near ties; tests report agreement with torch.cdist on every non-near-tie row.
//
// This is a reduction, not a contraction that is worth rounding to bf16: it runs on the fp32 FMA pipe.
// Work decomposition: unit = (128-row tile of z) x (split of the codebook); each CTA keeps its z tile in
// shared memory, streams 128-code tiles of the codebook through a cp.async double buffer and merges its
// per-row winner into a packed (ordered-score, index) 64-bit word with atomicMin, which is
// order-independent and therefore deterministic.  A second tiny kernel unpacks to int64.
#include "common.cuh"

namespace ivg {

constexpr int VQ_D = 64;          // embedding dim (vq_embed_dim = latent_channels = 64 in both configs)
constexpr int VQ_TN = 128;        // z rows per CTA
constexpr int VQ_TK = 128;        // codes per streamed tile
constexpr int VQ_PITCH = VQ_D + 4;  // floats; 272-byte rows -> conflict-free LDS.128 across 8 rows
constexpr int VQ_THREADS = 256;

__device__ __forceinline__ uint32_t ordered_bits(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void vq_enorm_kernel(const float* __restrict__ e, float* __restrict__ enorm, int K) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float4* row = reinterpret_cast<const float4*>(e + (size_t)k * VQ_D);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VQ_D / 4; ++i) {
    float4 v = __ldg(row + i);
    s = fmaf(v.x, v.x, s);
    s = fmaf(v.y, v.y, s);
    s = fmaf(v.z, v.z, s);
    s = fmaf(v.w, v.w, s);
  }
  enorm[k] = s;
}

__global__ void vq_fill_kernel(unsigned long long* packed, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) packed[i] = 0xFFFFFFFFFFFFFFFFull;
}

__global__ void vq_unpack_kernel(const unsigned long long* __restrict__ packed, long long* __restrict__ idx, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) idx[i] = (long long)(packed[i] & 0xFFFFFFFFull);
}

// thread (ty, tx): rows n = i*16 + ty, codes k = j*16 + tx (i, j = 0..7).  Lanes of a warp therefore read 16
// CONSECUTIVE codebook rows per LDS.128: row pitch 68 words -> bank offset 4*tx, conflict-free per quarter-warp
// (the first version used k = tx*4 + j: lane stride 272 words = 16 banks -> 8-way conflicts, ncu: 2.0e8 conflicts).
__global__ void __launch_bounds__(VQ_THREADS, 2)
vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ e, const float* __restrict__ enorm,
                 unsigned long long* __restrict__ packed, int N, int K, int tiles_per_split) {
  extern __shared__ __align__(16) float vq_smem[];
  float* zs = vq_smem;                          // [VQ_TN][VQ_PITCH]
  float* es = zs + VQ_TN * VQ_PITCH;            // [2][VQ_TK][VQ_PITCH]
  float* ns = es + 2 * VQ_TK * VQ_PITCH;        // [2][VQ_TK]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * VQ_TN;
  const int ktile0 = blockIdx.y * tiles_per_split;
  const int num_ktiles_total = (K + VQ_TK - 1) / VQ_TK;
  int ntiles = num_ktiles_total - ktile0;
  if (ntiles > tiles_per_split) ntiles = tiles_per_split;
  if (ntiles <= 0) return;

  // ---- stage z tile (rows past N are zero-filled; their results are never stored) ----
  for (int c = tid; c < VQ_TN * (VQ_D / 4); c += VQ_THREADS) {
    int r = c >> 4, q = c & 15;
    float* dst = zs + r * VQ_PITCH + q * 4;
    if (n0 + r < N) cp_async16(dst, z + (size_t)(n0 + r) * VQ_D + q * 4);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto load_tile = [&](int t, int buf) {
    const int k0 = (ktile0 + t) * VQ_TK;
    float* eb = es + buf * VQ_TK * VQ_PITCH;
    for (int c = tid; c < VQ_TK * (VQ_D / 4); c += VQ_THREADS) {
      int r = c >> 4, q = c & 15;
      float* dst = eb + r * VQ_PITCH + q * 4;
      if (k0 + r < K) cp_async16(dst, e + (size_t)(k0 + r) * VQ_D + q * 4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < VQ_TK) ns[buf * VQ_TK + tid] = (k0 + tid < K) ? __ldg(enorm + k0 + tid) : __int_as_float(0x7f800000);
  };
  load_tile(0, 0);
  cp_async_commit();

  float best[8];
  int bidx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = __int_as_float(0x7f800000); bidx[i] = 0x7fffffff; }

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_tile(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const float* eb = es + buf * VQ_TK * VQ_PITCH;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll 2
    for (int d = 0; d < VQ_D; d += 4) {
      float4 zv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int r = i * 16 + ty;
        zv[i] = *reinterpret_cast<const float4*>(zs + r * VQ_PITCH + d);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int c = j * 16 + tx;
        float4 ev = *reinterpret_cast<const float4*>(eb + c * VQ_PITCH + d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = acc[i][j];
          a = fmaf(zv[i].x, ev.x, a);
          a = fmaf(zv[i].y, ev.y, a);
          a = fmaf(zv[i].z, ev.z, a);
          a = fmaf(zv[i].w, ev.w, a);
          acc[i][j] = a;
        }
      }
    }
    // scores + running per-thread argmin (codes visited in increasing k within a thread: strict <)
    const int kbase = (ktile0 + t) * VQ_TK;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = j * 16 + tx;
      float en = ns[buf * VQ_TK + c];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = fmaf(-2.0f, acc[i][j], en);
        if (s < best[i]) { best[i] = s; bidx[i] = kbase + c; }
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled at t+2
  }

  // ---- merge across the 16 tx lanes that share a row (lanes differ in bits 0..3), then publish ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long p = ((unsigned long long)ordered_bits(best[i]) << 32) | (unsigned int)bidx[i];
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      unsigned long long q = __shfl_xor_sync(0xffffffffu, p, off);
      p = (q < p) ? q : p;
    }
    if (tx == 0) {
      int r = i * 16 + ty;
      if (n0 + r < N) atomicMin(packed + n0 + r, p);
    }
  }
}

// ---- packed-FMA variant (fma.rn.f32x2, sm_100+) ---------------------------------------------------------------
// Same decomposition, but every (row, code) pair keeps TWO partial dot products -- even and odd embedding dims --
// in one 64-bit register pair, so one FFMA2 retires two FMAs:   acc2 = fma2((z[d], z[d+1]), (e[d], e[d+1]), acc2).
// Definition (order 1, also in oracle/vq_argmin_ref.c):  dot = chain_even(d = 0,2,..,62) + chain_odd(d = 1,3,..,63),
// both chains sequential fp32 FMAs from +0; score / argmin as before.  Thread tile 8 rows x 4 codes.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

constexpr int VQ2_TK = 64;        // codes per streamed tile (16 tx x 4)
__global__ void __launch_bounds__(VQ_THREADS, 2)
vq_argmin_x2_kernel(const float* __restrict__ z, const float* __restrict__ e, const float* __restrict__ enorm,
                    unsigned long long* __restrict__ packed, int N, int K, int tiles_per_split) {
  extern __shared__ __align__(16) float vq_smem[];
  float* zs = vq_smem;                          // [VQ_TN][VQ_PITCH]
  float* es = zs + VQ_TN * VQ_PITCH;            // [2][VQ2_TK][VQ_PITCH]
  float* ns = es + 2 * VQ2_TK * VQ_PITCH;       // [2][VQ2_TK]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * VQ_TN;
  const int ktile0 = blockIdx.y * tiles_per_split;
  const int num_ktiles_total = (K + VQ2_TK - 1) / VQ2_TK;
  int ntiles = num_ktiles_total - ktile0;
  if (ntiles > tiles_per_split) ntiles = tiles_per_split;
  if (ntiles <= 0) return;
  for (int c = tid; c < VQ_TN * (VQ_D / 4); c += VQ_THREADS) {
    int r = c >> 4, q = c & 15;
    float* dst = zs + r * VQ_PITCH + q * 4;
    if (n0 + r < N) cp_async16(dst, z + (size_t)(n0 + r) * VQ_D + q * 4);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto load_tile = [&](int t, int buf) {
    const int k0 = (ktile0 + t) * VQ2_TK;
    float* eb = es + buf * VQ2_TK * VQ_PITCH;
    for (int c = tid; c < VQ2_TK * (VQ_D / 4); c += VQ_THREADS) {
      int r = c >> 4, q = c & 15;
      float* dst = eb + r * VQ_PITCH + q * 4;
      if (k0 + r < K) cp_async16(dst, e + (size_t)(k0 + r) * VQ_D + q * 4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < VQ2_TK) ns[buf * VQ2_TK + tid] = (k0 + tid < K) ? __ldg(enorm + k0 + tid) : __int_as_float(0x7f800000);
  };
  load_tile(0, 0);
  cp_async_commit();
  float best[8];
  int bidx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = __int_as_float(0x7f800000); bidx[i] = 0x7fffffff; }
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_tile(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* eb = es + buf * VQ2_TK * VQ_PITCH;
    unsigned long long acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
#pragma unroll 4
    for (int d = 0; d < VQ_D; d += 4) {
      unsigned long long za[8], zb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(zs + (i * 16 + ty) * VQ_PITCH + d);
        za[i] = pack2(v.x, v.y); zb[i] = pack2(v.z, v.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(eb + (j * 16 + tx) * VQ_PITCH + d);
        const unsigned long long ea = pack2(v.x, v.y), ebb = pack2(v.z, v.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) { fma2(acc[i][j], za[i], ea); fma2(acc[i][j], zb[i], ebb); }
      }
    }
    const int kbase = (ktile0 + t) * VQ2_TK;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = j * 16 + tx;
      const float en = ns[buf * VQ2_TK + c];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float s = fmaf(-2.0f, sum2(acc[i][j]), en);
        if (s < best[i]) { best[i] = s; bidx[i] = kbase + c; }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long p = ((unsigned long long)ordered_bits(best[i]) << 32) | (unsigned int)bidx[i];
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      unsigned long long q = __shfl_xor_sync(0xffffffffu, p, off);
      p = (q < p) ? q : p;
    }
    if (tx == 0) {
      int r = i * 16 + ty;
      if (n0 + r < N) atomicMin(packed + n0 + r, p);
    }
  }
}

static int vq_pick_splits(int N, int K, int num_sms) {
  int ntile = cdiv(N, VQ_TN), ktiles = cdiv(K, VQ_TK);
  int target = 6 * num_sms;  // ~3 waves at 2 CTAs/SM
  int splits = 1;
  while (ntile * splits < target && splits * 2 <= ktiles && (ktiles / (splits * 2)) >= 2) splits *= 2;
  return splits;
}

int g_vq_order = 0;   // 0: single sequential FMA chain (FFMA kernel); 1: even/odd chains (packed FFMA2 kernel)

int vq_argmin_launch(const float* z, const float* e, float* enorm_ws, unsigned long long* packed_ws, long long* idx,
                     int N, int K, int D, int num_sms, cudaStream_t stream) {
  IVG_CHECK(D == VQ_D, "vq_argmin: D must be %d, got %d", VQ_D, D);
  IVG_CHECK(N >= 0 && K > 0 && K < 0x7fffffff, "vq_argmin: bad N=%d K=%d", N, K);
  if (N == 0) return 0;
  IVG_CHECK(((uintptr_t)z & 15) == 0 && ((uintptr_t)e & 15) == 0, "vq_argmin: z/e must be 16-byte aligned");
  static PerDeviceOnce attr_once;
  const size_t smem = (size_t)(VQ_TN * VQ_PITCH + 2 * VQ_TK * VQ_PITCH + 2 * VQ_TK) * sizeof(float);
  const size_t smem2 = (size_t)(VQ_TN * VQ_PITCH + 2 * VQ2_TK * VQ_PITCH + 2 * VQ2_TK) * sizeof(float);
  if (attr_once.pending()) {
    IVG_CUDA(cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IVG_CUDA(cudaFuncSetAttribute(vq_argmin_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    attr_once.mark();
  }
  vq_enorm_kernel<<<cdiv(K, 256), 256, 0, stream>>>(e, enorm_ws, K);
  vq_fill_kernel<<<cdiv(N, 256), 256, 0, stream>>>(packed_ws, N);
  if (g_vq_order == 0) {
    int splits = vq_pick_splits(N, K, num_sms);
    int ktiles = cdiv(K, VQ_TK);
    int tiles_per_split = cdiv(ktiles, splits);
    dim3 grid(cdiv(N, VQ_TN), cdiv(ktiles, tiles_per_split));
    vq_argmin_kernel<<<grid, VQ_THREADS, smem, stream>>>(z, e, enorm_ws, packed_ws, N, K, tiles_per_split);
  } else {
    int ktiles = cdiv(K, VQ2_TK);
    int ntile = cdiv(N, VQ_TN), splits = 1;
    while (ntile * splits < 6 * num_sms && splits * 2 <= ktiles && (ktiles / (splits * 2)) >= 2) splits *= 2;
    int tiles_per_split = cdiv(ktiles, splits);
    dim3 grid(ntile, cdiv(ktiles, tiles_per_split));
    vq_argmin_x2_kernel<<<grid, VQ_THREADS, smem2, stream>>>(z, e, enorm_ws, packed_ws, N, K, tiles_per_split);
  }
  vq_unpack_kernel<<<cdiv(N, 256), 256, 0, stream>>>(packed_ws, idx, N);
  count_launch(4);
  IVG_LAUNCH_CHECK();
  return 0;
}

}  // namespace ivg
