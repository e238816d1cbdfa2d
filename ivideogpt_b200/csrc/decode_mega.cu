// Persistent decode megakernel: the whole per-token rollout loop of HF `generate` (reference
// inference/predict.py:64-69 -> transformers GenerationMixin: 237 x [12/24 decoder layers -> lm_head -> top-k
// sample]) as ONE cooperative launch.  One CTA per SM stays resident for all steps; phases are separated by a
// device-wide barrier (monotonic counter in global memory) instead of ~100 kernel launches per token.
//
// Why: a decode step at batch 64 moves ~250 MB of weights + the KV cache and does 16 GFLOP -- HBM-bound by two
// orders of magnitude over its launch overhead.  As separate kernels the step cost 1.6 ms (13 us of fixed cost per
// skinny GEMM); the roofline (weights + KV at 6.5 TB/s) is ~0.27 ms.
//
// Per step:  embed+norm | { qkv GEMM | RoPE+append+attention | o GEMM (split-K) | add+norm | gate/up GEMM+SwiGLU |
//            down GEMM (split-K) | add+norm } x layers | lm_head GEMM | sample+append
// GEMM phases (tcgen05, M = 128-row UMMA with the first B rows valid, N = 16 per work item so every SM streams its
// own slice of the weight matrix):
//   * B operand (weights) : TMA, SWIZZLE_128B, whole K-slab of the item in one burst; the slab of the NEXT phase is
//     issued before the device-wide barrier, so the weight fetch overlaps the barrier + the neighbour phase.
//   * A operand (activations, <= 128 x 1024 bf16): written by the previous phase with ordinary stores, so it is
//     loaded with ordinary loads and laid out in shared memory by hand in the 128B-swizzled K-major format
//     (16-byte chunk c of row r lands at chunk c ^ (r & 7)); rows >= B of the UMMA tile alias whatever follows in
//     shared memory -- their TMEM lanes are never read.
//   * split-K partials go to a [3][B][hidden] fp32 buffer and are summed in a FIXED order by the following add+norm
//     phase: no atomics, bit-reproducible.
// bf16 only (the fp32/TF32 parity path keeps the multi-kernel CUDA-graph step).
#include <cooperative_groups.h>

#include "common.cuh"
#include "decode_mega.cuh"

namespace ivg {

struct MegaSmem {
  uint8_t* a;            // A region (1024-aligned)
  uint8_t* b[2];         // weight slabs
  uint64_t* bfull;       // [2]
  uint64_t* mma_done;    // [1]
  uint32_t* tmem_holder;
};

struct MegaCtx {
  MegaSmem sm;
  uint32_t tmem_base;
  uint32_t epoch;          // barrier target (thread 0)
  // Weight-slab queue (thread 0 only).  This CTA's slab loads form one deterministic sequence over the whole kernel;
  // load #i always goes to buffer i & 1 and is consumed with parity (i >> 1) & 1.  At most two are in flight.
  uint32_t issued, consumed;
  int phase_issued;        // items of the CURRENT/coming phase whose slab has been issued already
  uint32_t mphase;         // parity of mma_done (all threads)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// device-wide barrier; returns false when it timed out (error flag is set, every CTA leaves the kernel)
__device__ __forceinline__ bool grid_barrier(const MegaParams& p, MegaCtx& c) {
  __syncthreads();
  if (threadIdx.x == 0) {
    c.epoch += gridDim.x;
    __threadfence();
    atomicAdd(p.barrier, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(p.barrier) < c.epoch) {
      if (clock64() - t0 > (1ll << 32)) { *p.error = 1; break; }   // ~2 s: never hang the GPU
      if (*reinterpret_cast<volatile int*>(p.error)) break;
    }
    __threadfence();
  }
  __syncthreads();
  return *reinterpret_cast<volatile int*>(p.error) == 0;
}

// ---- weight slab TMA: item -> (n0, k0, Kc); Kc/64 boxes of {64 k, 16 rows} land on bfull[buffer] ----
__device__ __forceinline__ void issue_slab(MegaCtx& c, const CUtensorMap* map, int n0, int k0, int Kc) {
  const int buf = (int)(c.issued & 1u);
  ++c.issued;
  const int nkb = Kc / 64;
  mbar_expect_tx(c.sm.bfull + buf, (uint32_t)(nkb * MEGA_BN * 128));
  for (int j = 0; j < nkb; ++j)
    tma_load_2d(c.sm.b[buf] + j * (MEGA_BN * 128), map, c.sm.bfull + buf, k0 + j * 64, n0);
}

// ---- A operand: rows [0, B) x k [k0, k0+Kc) of a row-major bf16 matrix -> swizzled 64-row K-major tiles ----
// tile j (k-block j) occupies a_tile_bytes = a_rows*128 bytes; a_rows = 64 when B <= 64 else 128.
__device__ __forceinline__ void load_a(const MegaParams& p, MegaCtx& c, const __nv_bfloat16* A, long long lda, int k0,
                                       int Kc, int a_rows) {
  const int nkb = Kc / 64;
  const int chunks = nkb * a_rows * 8;            // 16-byte chunks
  for (int i = threadIdx.x; i < chunks; i += MEGA_THREADS) {
    const int ch = i & 7;
    const int r = (i >> 3) % a_rows;
    const int j = (i >> 3) / a_rows;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < p.B) v = *reinterpret_cast<const uint4*>(A + (size_t)r * lda + k0 + j * 64 + ch * 8);
    uint8_t* dst = c.sm.a + (size_t)j * a_rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = v;
  }
  fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core's async proxy
  __syncthreads();
}

enum { EPI_STORE_BF16 = 0, EPI_PARTIAL_F32 = 1, EPI_SWIGLU = 2, EPI_LOGITS = 3 };

struct GemmPhase {
  const CUtensorMap* map;   // weights [N, K]
  int N, K, ksplits;        // work items = ceil(N/16) * ksplits, item K = K / ksplits
  const __nv_bfloat16* A;
  long long lda;
  int epi;
  void* out;                // bf16 / fp32 destination
  long long ldo;
};

__device__ __forceinline__ int phase_items(const GemmPhase& g) { return ((g.N + MEGA_BN - 1) / MEGA_BN) * g.ksplits; }

// issue this CTA's slabs of phase g up to item index `upto` (exclusive), in order
__device__ __forceinline__ void issue_items(MegaCtx& c, const GemmPhase& g, int upto) {
  const int items = phase_items(g);
  const int ntiles = items / g.ksplits;
  const int Kc = g.K / g.ksplits;
  while (c.phase_issued < upto) {
    const int w = blockIdx.x + c.phase_issued * gridDim.x;
    if (w >= items) break;
    issue_slab(c, g.map, (w % ntiles) * MEGA_BN, (w / ntiles) * Kc, Kc);
    ++c.phase_issued;
  }
}

// called before the barrier that precedes phase g: weights do not depend on anything, start fetching them now
__device__ __forceinline__ void prefetch_phase(MegaCtx& c, const GemmPhase& g) {
  if (threadIdx.x != 0) return;
  c.phase_issued = 0;
  issue_items(c, g, 1);
}

__device__ void gemm_phase(const MegaParams& p, MegaCtx& c, const GemmPhase& g) {
  const int items = phase_items(g);
  const int ntiles = items / g.ksplits;
  const int Kc = g.K / g.ksplits;
  const int nkb = Kc / 64;
  const int a_rows = p.B <= 64 ? 64 : 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t IDESC = umma_idesc(1, 128, MEGA_BN);
  int loaded_split = -1;
  int it = 0;
  for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
    const int tile = w % ntiles, split = w / ntiles;
    if (split != loaded_split) {          // (re)load the activation slab for this K range
      load_a(p, c, g.A, g.lda, split * Kc, Kc, a_rows);
      loaded_split = split;
    }
    if (threadIdx.x == 0) {
      issue_items(c, g, it + 2);          // this item (if not prefetched) and the next one (other buffer)
      const int buf = (int)(c.consumed & 1u);
      const uint32_t par = (c.consumed >> 1) & 1u;
      ++c.consumed;
      mbar_wait(c.sm.bfull + buf, par);
      tc_fence_after();
      const uint32_t a0 = smem_u32(c.sm.a), b0 = smem_u32(c.sm.b[buf]);
      for (int j = 0; j < nkb; ++j) {
        const uint64_t adesc = umma_desc_sw128_kmajor(a0 + (uint32_t)(j * a_rows * 128));
        const uint64_t bdesc = umma_desc_sw128_kmajor(b0 + (uint32_t)(j * MEGA_BN * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss<false>(c.tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (j | k) ? 1u : 0u);
      }
      umma_commit(c.sm.mma_done);
    }
    // ---- epilogue: warps 4..7 own TMEM lane quadrants 0..3 ----
    if (warp >= 4) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      mbar_wait(c.sm.mma_done, c.mphase);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x32b_x16(c.tmem_base + ((uint32_t)(q * 32) << 16), r);
      tmem_ld_wait();
      if (row < p.B) {
        const int n0 = tile * MEGA_BN;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (g.epi == EPI_STORE_BF16) {
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)row * g.ldo + n0);
          op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        } else if (g.epi == EPI_PARTIAL_F32) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + ((size_t)split * p.B + row) * g.ldo + n0);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else if (g.epi == EPI_SWIGLU) {
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)row * g.ldo + (n0 >> 1));
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = silu_f(v[2 * i]) * v[2 * i + 1];
          op[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        } else {
          float* op = reinterpret_cast<float*>(g.out) + (size_t)row * g.ldo + n0;
#pragma unroll
          for (int i = 0; i < 16; ++i) if (n0 + i < g.N) op[i] = v[i];
        }
      }
      tc_fence_before();
    }
    c.mphase ^= 1;          // every thread tracks the parity (only warps 4..7 wait on it)
    __syncthreads();        // MMA retired (epilogue observed mma_done): TMEM accumulator, A slab and this weight
                            // buffer may be reused
  }
}

// ---- add split-K partials (fixed order) + RMSNorm -> xn ; or embedding gather + RMSNorm ----
__device__ void norm_phase(const MegaParams& p, const float* w, int nparts, const long long* tok_row0, int tok_col) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (MEGA_THREADS / 32) + warp;
  const int H = p.hidden;
  for (int m = gw; m < p.B; m += gridDim.x * (MEGA_THREADS / 32)) {
    float* xr = p.x + (size_t)m * H;
    const float* src = xr;
    if (tok_row0) {
      long long id = tok_row0[(size_t)m * p.tok_stride + tok_col];
      id = id < 0 ? 0 : (id >= p.vocab ? p.vocab - 1 : id);
      src = p.embed + (size_t)id * H;
    }
    float ss = 0.f;
    for (int i = lane * 4; i < H; i += 128) {
      float4 v = *reinterpret_cast<const float4*>(src + i);
      for (int s = 0; s < nparts; ++s) {
        const float4 q = *reinterpret_cast<const float4*>(p.part + ((size_t)s * p.B + m) * H + i);
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      *reinterpret_cast<float4*>(xr + i) = v;
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float r = rsqrtf(ss / (float)H + p.eps);
    __nv_bfloat16* yr = p.xn + (size_t)m * H;
    for (int i = lane * 4; i < H; i += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(w + i));
      uint2 o;
      o.x = pack_bf16x2(g.x * (v.x * r), g.y * (v.y * r));
      o.y = pack_bf16x2(g.z * (v.z * r), g.w * (v.w * r));
      *reinterpret_cast<uint2*>(yr + i) = o;
    }
  }
}

// ---- RoPE + KV append + attention over the cache for one (b, head); same arithmetic as decode_attn_fused_kernel ----
__device__ void attention_item(const MegaParams& p, int layer, int bh, int pos, float* smem_f) {
  const int heads = p.heads, Lmax = p.Lmax, Hd = p.hidden;
  const int Lcur = pos + 1;
  float* sc = smem_f;
  float* qs = smem_f + Lmax + 8;
  float* red = qs + 64;
  const int b = bh / heads, hh = bh - b * heads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __nv_bfloat16* kslab = p.kcache + ((size_t)layer * p.B * heads + bh) * Lmax * 64;
  __nv_bfloat16* vslab = p.vcache + ((size_t)layer * p.B * heads + bh) * 64 * Lmax;
  if (tid < 32) {
    const __nv_bfloat16* row = p.qkv + (size_t)b * 3 * Hd;
    const float cs = __ldg(p.cos_tab + (size_t)pos * 32 + tid), sn = __ldg(p.sin_tab + (size_t)pos * 32 + tid);
    const float q0 = __bfloat162float(row[hh * 64 + tid]), q1 = __bfloat162float(row[hh * 64 + tid + 32]);
    const float k0 = __bfloat162float(row[Hd + hh * 64 + tid]), k1 = __bfloat162float(row[Hd + hh * 64 + tid + 32]);
    qs[tid] = __bfloat162float(__float2bfloat16_rn(q0 * cs - q1 * sn)) * 0.125f;
    qs[tid + 32] = __bfloat162float(__float2bfloat16_rn(q1 * cs + q0 * sn)) * 0.125f;
    kslab[(size_t)pos * 64 + tid] = __float2bfloat16_rn(k0 * cs - k1 * sn);
    kslab[(size_t)pos * 64 + tid + 32] = __float2bfloat16_rn(k1 * cs + k0 * sn);
    vslab[(size_t)tid * Lmax + pos] = row[2 * Hd + hh * 64 + tid];
    vslab[(size_t)(tid + 32) * Lmax + pos] = row[2 * Hd + hh * 64 + tid + 32];
  }
  __syncthreads();
  const int sub = tid & 7, rslot = tid >> 3;
  float qreg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qreg[i] = qs[sub * 8 + i];
  for (int l0 = 0; l0 < Lcur; l0 += 32) {
    const int l = l0 + rslot;
    float part = 0.f;
    if (l < Lcur) {
      const uint4 u = *reinterpret_cast<const uint4*>(kslab + (size_t)l * 64 + sub * 8);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        part = fmaf(qreg[2 * i], f.x, part);
        part = fmaf(qreg[2 * i + 1], f.y, part);
      }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    if (sub == 0 && l < Lcur) sc[l] = part;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int l = tid; l < Lcur; l += MEGA_THREADS) mx = fmaxf(mx, sc[l]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  float sum = 0.f;
  for (int l = tid; l < Lcur + 8; l += MEGA_THREADS) {
    const float e = l < Lcur ? __expf(sc[l] - mx) : 0.f;
    sc[l] = e;
    sum += e;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) red[8 + warp] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[8 + i];
  const float inv = 1.0f / tot;
  for (int d = warp; d < 64; d += 8) {
    const __nv_bfloat16* vr = vslab + (size_t)d * Lmax;
    float a = 0.f;
    for (int l = lane * 8; l < Lcur; l += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(vr + l);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        a += (l + 2 * i < Lcur) ? sc[l + 2 * i] * f.x : 0.f;
        a += (l + 2 * i + 1 < Lcur) ? sc[l + 2 * i + 1] * f.y : 0.f;
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) p.ao[(size_t)b * Hd + hh * 64 + d] = __float2bfloat16_rn(a * inv);
  }
  __syncthreads();   // smem scratch reused by the next item
}

// ---- sampling of one logits row by one CTA (argmax, or top-k radix select + inverse CDF as topk_sample_kernel) ----
__device__ __forceinline__ uint32_t mega_fkey(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float mega_uniform(unsigned long long seed, unsigned long long step, unsigned long long row) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (step * 0x100000001B3ull + row + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)((z >> 40) + 0.5) * (1.0f / 16777216.0f);
}

__device__ void sample_row(const MegaParams& p, int b, int pos, uint32_t* smem_u) {
  const float* row = p.logits + (size_t)b * p.ldl;
  const int V = p.vocab;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long* out = p.tokens + (size_t)b * p.tok_stride + pos + 1;
  __shared__ float s_redf[8];
  __shared__ int s_redi[8];
  if (!p.do_sample) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int c = tid; c < V; c += MEGA_THREADS) { const float v = row[c]; if (v > bv || (v == bv && c < bi)) { bv = v; bi = c; } }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_redf[warp] = bv; s_redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < 8; ++i) if (s_redf[i] > bv || (s_redf[i] == bv && s_redi[i] < bi)) { bv = s_redf[i]; bi = s_redi[i]; }
      *out = bi;
    }
    __syncthreads();
    return;
  }
  uint32_t* keys = smem_u;
  uint32_t* hist = smem_u + V;
  __shared__ uint32_t s_prefix, s_remaining;
  __shared__ float s_hi[MEGA_THREADS];
  __shared__ int s_win, s_lastmass;
  for (int c = tid; c < V; c += MEGA_THREADS) keys[c] = mega_fkey(row[c] * p.inv_temp);
  if (tid == 0) { s_prefix = 0; s_remaining = (uint32_t)(p.topk < V ? p.topk : V); s_win = 0x7fffffff; s_lastmass = 0; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += MEGA_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int c = tid; c < V; c += MEGA_THREADS) {
      const uint32_t kk = keys[c];
      if ((kk & mask) == prefix) atomicAdd(&hist[(kk >> shift) & 255], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t rem = s_remaining;
      int bb = 255;
      for (; bb > 0; --bb) { if (hist[bb] >= rem) break; rem -= hist[bb]; }
      s_prefix = prefix | ((uint32_t)bb << shift);
      s_remaining = rem;
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;
  uint32_t mk = 0;
  for (int c = tid; c < V; c += MEGA_THREADS) mk = max(mk, keys[c]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mk = max(mk, __shfl_xor_sync(0xffffffffu, mk, off));
  if (lane == 0) s_redi[warp] = (int)mk;
  __syncthreads();
  for (int i = 0; i < 8; ++i) mk = max(mk, (uint32_t)s_redi[i]);
  const float mx = __uint_as_float((mk & 0x80000000u) ? (mk ^ 0x80000000u) : ~mk);
  const int cpt = (V + MEGA_THREADS - 1) / MEGA_THREADS;
  const int c0 = tid * cpt, c1 = min(V, c0 + cpt);
  float local = 0.f;
  for (int c = c0; c < c1; ++c) {
    const uint32_t kk = keys[c];
    if (kk >= kth) local += __expf(__uint_as_float((kk & 0x80000000u) ? (kk ^ 0x80000000u) : ~kk) - mx);
  }
  float incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) s_redf[warp] = incl;
  __syncthreads();
  float wbase = 0.f, total = 0.f;
  for (int i = 0; i < 8; ++i) { const float v = s_redf[i]; if (i < warp) wbase += v; total += v; }
  const float hi = wbase + incl;
  s_hi[tid] = hi;
  const unsigned long long seed = p.dseed ? *p.dseed : 0ull;
  const float u = mega_uniform(seed, (unsigned long long)pos, (unsigned long long)b) * total;
  if (local > 0.f) {
    atomicMax(&s_lastmass, tid);
    if (hi >= u) atomicMin(&s_win, tid);
  }
  __syncthreads();
  const int win = (s_win == 0x7fffffff) ? s_lastmass : s_win;
  if (tid == win) {
    float acc = win > 0 ? s_hi[win - 1] : 0.f;
    int pick = -1, last = c0;
    for (int c = c0; c < c1; ++c) {
      const uint32_t kk = keys[c];
      if (kk >= kth) {
        acc += __expf(__uint_as_float((kk & 0x80000000u) ? (kk ^ 0x80000000u) : ~kk) - mx);
        last = c;
        if (acc >= u) { pick = c; break; }
      }
    }
    *out = pick >= 0 ? pick : last;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(const MegaParams p) {
  extern __shared__ uint8_t mega_raw[];
  MegaCtx c;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mega_raw) + 1023) & ~(uintptr_t)1023);
  c.sm.a = base;
  c.sm.b[0] = base + MEGA_A_BYTES;
  c.sm.b[1] = c.sm.b[0] + MEGA_B_BYTES;
  c.sm.bfull = reinterpret_cast<uint64_t*>(c.sm.b[1] + MEGA_B_BYTES);
  c.sm.mma_done = c.sm.bfull + 2;
  c.sm.tmem_holder = reinterpret_cast<uint32_t*>(c.sm.mma_done + 1);
  c.epoch = 0; c.issued = 0; c.consumed = 0; c.phase_issued = 0; c.mphase = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(c.sm.bfull, 1); mbar_init(c.sm.bfull + 1, 1); mbar_init(c.sm.mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(c.sm.tmem_holder, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *c.sm.tmem_holder;

  const int H = p.hidden;
  const int pos0 = *p.dpos;
  float* smem_f = reinterpret_cast<float*>(c.sm.a);
  uint32_t* smem_u = reinterpret_cast<uint32_t*>(c.sm.a);
  bool ok = true;

  for (int step = 0; step < p.steps && ok; ++step) {
    const int pos = pos0 + step;
    GemmPhase qkv_g{&p.lw[0].wqkv, 3 * H, H, 1, p.xn, H, EPI_STORE_BF16, p.qkv, 3 * H};
    prefetch_phase(c, qkv_g);
    norm_phase(p, p.lw[0].n1, 0, p.tokens, pos);                       // x = E[token], xn = rmsnorm(x)
    if (!(ok = grid_barrier(p, c))) break;
    for (int l = 0; l < p.layers && ok; ++l) {
      const MegaLayer& L = p.lw[l];
      qkv_g.map = &L.wqkv;
      gemm_phase(p, c, qkv_g);
      GemmPhase o_g{&L.wo, H, H, p.o_splits, p.ao, H, EPI_PARTIAL_F32, p.part, H};
      prefetch_phase(c, o_g);
      if (!(ok = grid_barrier(p, c))) break;
      for (int bh = blockIdx.x; bh < p.B * p.heads; bh += gridDim.x) attention_item(p, l, bh, pos, smem_f);
      if (!(ok = grid_barrier(p, c))) break;
      gemm_phase(p, c, o_g);
      GemmPhase gu_g{&L.wgu, 2 * p.inter, H, 1, p.xn, H, EPI_SWIGLU, p.act, p.inter};
      prefetch_phase(c, gu_g);
      if (!(ok = grid_barrier(p, c))) break;
      norm_phase(p, L.n2, p.o_splits, nullptr, 0);                    // x += o partials; xn = rmsnorm(x) * n2
      if (!(ok = grid_barrier(p, c))) break;
      gemm_phase(p, c, gu_g);
      GemmPhase d_g{&L.wd, H, p.inter, p.d_splits, p.act, p.inter, EPI_PARTIAL_F32, p.part, H};
      prefetch_phase(c, d_g);
      if (!(ok = grid_barrier(p, c))) break;
      gemm_phase(p, c, d_g);
      const bool last = (l == p.layers - 1);
      GemmPhase nx_g{last ? p.lm_head : &p.lw[l + 1].wqkv, last ? p.vocab : 3 * H, H, 1, p.xn, H,
                     last ? EPI_LOGITS : EPI_STORE_BF16, last ? (void*)p.logits : (void*)p.qkv,
                     last ? p.ldl : (long long)(3 * H)};
      prefetch_phase(c, nx_g);
      if (!(ok = grid_barrier(p, c))) break;
      norm_phase(p, last ? p.norm_f : p.lw[l + 1].n1, p.d_splits, nullptr, 0);   // x += down partials; next norm
      if (!(ok = grid_barrier(p, c))) break;
      if (last) {
        gemm_phase(p, c, nx_g);                                       // lm_head
        if (!(ok = grid_barrier(p, c))) break;
        for (int b = blockIdx.x; b < p.B; b += gridDim.x) sample_row(p, b, pos, smem_u);
        if (!(ok = grid_barrier(p, c))) break;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && ok) *p.dpos = pos0 + p.steps;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(c.tmem_base, 32); }
}

int decode_mega_launch(const MegaParams& p, int num_sms, cudaStream_t st) {
  IVG_CHECK(p.B >= 1 && p.B <= 128, "decode_mega: batch %d not in [1,128]", p.B);
  IVG_CHECK(p.hidden % 64 == 0 && p.hidden <= MEGA_MAXK, "decode_mega: hidden %d unsupported", p.hidden);
  IVG_CHECK(p.o_splits >= 1 && p.o_splits <= MEGA_MAX_SPLITS && p.hidden % (64 * p.o_splits) == 0,
            "decode_mega: bad o_splits %d for hidden %d", p.o_splits, p.hidden);
  IVG_CHECK(p.d_splits >= 1 && p.d_splits <= MEGA_MAX_SPLITS && p.inter % (64 * p.d_splits) == 0 &&
                p.inter / p.d_splits <= MEGA_MAXK,
            "decode_mega: bad d_splits %d for intermediate size %d", p.d_splits, p.inter);
  IVG_CHECK(p.hidden == p.heads * 64, "decode_mega: head_dim must be 64");
  const int a_rows = p.B <= 64 ? 64 : 128;
  IVG_CHECK((long long)a_rows * p.hidden * 2 <= MEGA_A_BYTES && (long long)a_rows * (p.inter / p.d_splits) * 2 <= MEGA_A_BYTES,
            "decode_mega: batch %d with K %d does not fit the shared-memory activation slab", p.B, p.hidden);
  IVG_CHECK((size_t)(p.Lmax + 8 + 64 + 16) * 4 <= MEGA_A_BYTES && (size_t)(p.vocab + 256) * 4 <= MEGA_A_BYTES,
            "decode_mega: Lmax/vocab too large for the scratch region");
  IVG_CHECK(p.Lmax % 8 == 0, "decode_mega: Lmax must be a multiple of 8");
  static bool attr_set = false;
  if (!attr_set) {
    IVG_CUDA(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_SMEM));
    attr_set = true;
  }
  int max_blocks = 0;
  IVG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, decode_mega_kernel, MEGA_THREADS, MEGA_SMEM));
  IVG_CHECK(max_blocks >= 1, "decode_mega: kernel does not fit on an SM");
  void* args[] = {const_cast<MegaParams*>(&p)};
  IVG_CUDA(cudaLaunchCooperativeKernel((const void*)decode_mega_kernel, dim3(num_sms), dim3(MEGA_THREADS), args,
                                       (size_t)MEGA_SMEM, st));
  count_launch();
  return 0;
}

}  // namespace ivg
