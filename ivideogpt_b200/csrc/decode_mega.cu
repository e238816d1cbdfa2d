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
//   * B operand (weights) : pre-packed once (ivgpt_mega_pack_weight) into the exact 128B-swizzled shared-memory image
//     of every 16-row work item, so an item's whole K-slab is ONE contiguous 1-D bulk copy (8-32 KB); up to four slabs
//     are in flight per CTA and the first slabs of the NEXT phase are issued before the device-wide barrier, so the
//     weight stream overlaps the barrier + the neighbour phase.  (Until v6 the slab was 12 tensor-map boxes of 16 x
//     128 B with one slab of look-ahead: ~3.4 us per item, 1.4 TB/s on the lm_head phase.)
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
  uint8_t* b0;           // weight slab buffers: nbuf x slab_bytes, directly above the activation slab
  uint32_t slab_bytes;   // 16 rows x max K per item x 2 B
  uint32_t nbuf;         // 2..4
  uint32_t a_bytes;      // activation slab / attention ring / sampler scratch region size
  uint64_t* bfull;       // [4]
  uint64_t* mma_done;    // [1]
  uint64_t* abar;        // [1] activation slab landed (a_bulk)
  uint64_t* bempty;      // [4] gemm_mode 1: the MMAs that read weight slab buffer i have retired
  uint32_t* tmem_holder;
  uint64_t* ring_bar;    // [8 warps][8 slots] attention ring: slot filled
  float* sc;             // attention scratch, MEGA_SC_BYTES
};

struct MegaCtx {
  MegaSmem sm;
  uint32_t tmem_base;
  uint32_t epoch;          // barrier target (thread 0)
  // Weight-slab queue (thread 0 only).  This CTA's slab loads form one deterministic sequence over the whole kernel;
  // load #i always goes to buffer i % nbuf and is consumed with parity (i / nbuf) & 1.  At most nbuf are in flight.
  uint32_t issued, consumed;
  int phase_issued;        // items of the CURRENT/coming phase whose slab has been issued already
  int phase_consumed;      // gemm_mode 0: items of the current phase whose slab has been waited for
  uint32_t wait_par;       // gemm_mode 0: expected parity of bfull[i] in bit i (buffers are re-numbered per phase)
  uint32_t mphase;         // parity of mma_done (all threads)
  uint32_t aphase;         // parity of abar (thread 0)
};

// 16-byte global load cached in L2 only.  With ~200 KB of the SM's 228 KB configured as shared memory the L1 data
// array is only ~29 KB; default (L1-allocating) loads could keep no more than that in flight per SM, which capped the
// attention phase at ~15 GB/s per SM (2.2 TB/s chip-wide, ncu/phase counters of v2/v3).  .cg loads do not allocate.
__device__ __forceinline__ uint4 ldg_cg_v4(const void* p) {
  uint4 v;
  asm("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- activation images ----
// The GEMM A operands produced inside the kernel (xn, ao, act) live in global memory ALREADY in the
// 128B-swizzled K-major shared-memory image the tensor core reads: [k-block = k/64][row][chunk ^ (row & 7)][8 bf16],
// a_rows (64 or 128) rows per k-block.  A K range of a phase is then one contiguous run of k-blocks: the whole
// activation slab is ONE 1-D bulk copy by the copy engine instead of 24 cp.async per thread + wait + __syncthreads.
__device__ __forceinline__ size_t a_off(const MegaParams& p, int m, int k, long long /*ld: row-major stride, unused*/) {
  const int a_rows = p.a_rows;
  return (size_t)(k >> 6) * (size_t)(a_rows * 64) + (size_t)m * 64 + (size_t)((((k >> 3) & 7) ^ (m & 7)) << 3) + (size_t)(k & 7);
}

// device-wide barrier; returns false when it timed out (error flag is set, every CTA leaves the kernel).
// Arrival is a release-reduction, the wait an acquire-load: no full sc fences (v1 used __threadfence() on both sides
// and cost ~2.8 us per barrier, 87 barriers per decode step).
template <bool FENCE>
__device__ __forceinline__ bool grid_barrier(const MegaParams& p, MegaCtx& c) {
  __shared__ int s_ok;
  // FENCE: this CTA's generic-proxy writes (activation images in global memory, scratch in the shared activation region)
  // must be ordered before the copy-engine (async proxy) accesses that follow the barrier, here and in other CTAs.  Only the
  // barriers that follow a phase producing an activation image (norm -> xn, attention -> ao, gate/up -> act) need it.
#ifdef IVG_MEGA_FENCE_ALL
  asm volatile("fence.proxy.async;\n" ::: "memory");
#else
  if constexpr (FENCE) asm volatile("fence.proxy.async;\n" ::: "memory");
#endif
  __syncthreads();
  if (threadIdx.x == 0) {
    c.epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.barrier) : "memory");
    const long long t0 = clock64();
    int ok = 1;
    while (ld_acquire_u32(p.barrier) < c.epoch) {
      if (clock64() - t0 > (1ll << 32)) { *p.error = 1; ok = 0; break; }   // ~2 s: never hang the GPU
    }
    // (v3 had every thread read the error flag with a volatile system-scope load after each barrier: ~1 us x 87
    //  barriers per step, visible as LDG.E.STRONG.SYS in the ncu source view.)
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// 1-D bulk copy (TMA engine, no tensor map) global -> shared, completing on an mbarrier.  dst/src 16-byte aligned,
// bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// mbarrier wait that can never hang the GPU: after ~1 s the error flag is raised (the host raises) and the wait falls through
__device__ __forceinline__ void mbar_wait_bounded(const MegaParams& p, uint64_t* bar, uint32_t parity, int code) {
#ifndef IVG_MEGA_BOUNDED_WAITS
  // default: plain spin.  The bounded form (debug builds, -DIVG_MEGA_BOUNDED_WAITS) costs ~1.6 ms per 236-step rollout in
  // register pressure (same-box A/B, profiles/r01/mega_build_variants_ab.txt)
  mbar_wait(bar, parity);
  return;
#endif
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > (1ll << 31)) { *p.error = code; break; }
  }
}

// ---- weight slab: item (tile, split) of a packed matrix -> one contiguous copy of Kc * 32 bytes ----
// packed layout (ivgpt_mega_pack_weight): [tile = n / 16][k-block = k / 64][row = n % 16][chunk ^ (row & 7)][8 bf16]
// A phase has its own tile width bn (16 weight rows per item where one round over the SMs covers the matrix; wider for
// gate/up and lm_head, whose 16-row items needed 3 and 7 rounds of 48 MMA issues each) and with it its own slab size,
// buffer count and buffer placement: buffer i of the phase sits at slab_off + i * slab_bytes of the operand area, load #j of
// the phase goes to buffer j % nbuf.  Phases never overlap in time, so the buffers of one phase may alias another's.
struct GemmPhase;
__device__ __forceinline__ void issue_slab(MegaCtx& c, const GemmPhase& g, int tile, int k0, int Kc);

// (Until round-1 v10 the activation slab was row-major in global memory and laid out in shared memory by 24 cp.async per
// thread + wait + __syncthreads: ~2 us per phase, 97 us per step; the bulk copy of a pre-swizzled image costs 24.)
enum { EPI_STORE_BF16 = 0, EPI_PARTIAL_F32 = 1, EPI_SWIGLU = 2, EPI_LOGITS = 3 };
// fused norms: EPI_PARTIAL_F32 phases carry `resid_cnt` (the tile's last arriver adds the partials to x and writes bf16(x));
// phases whose A operand is that image carry `norm_a` (epilogue scales row b by rstd_b computed from the shared-memory slab)

struct GemmPhase {
  const __nv_bfloat16* w;   // packed weights of [N, K]
  int N, K, ksplits;        // work items = ceil(N/bn) * ksplits, item K = K / ksplits
  const __nv_bfloat16* A;
  long long lda;
  int epi;
  void* out;                // bf16 / fp32 destination
  long long ldo;
  // gemm_mode 0 slab geometry (filled by phase_geometry); gemm_mode 1 ignores it
  int bn;                   // weight rows per work item (multiple of 16)
  uint32_t slab_off, slab_bytes, nbuf;
  int norm_a;               // fused norms: A is bf16(x); multiply row b of the result by rsqrt(mean(x_b^2) + eps)
  unsigned int* resid_cnt;  // fused norms: per-tile arrival counters of a split-K projection whose result is added to x
};

__device__ __forceinline__ int phase_items(const GemmPhase& g) { return ((g.N + g.bn - 1) / g.bn) * g.ksplits; }

__device__ __forceinline__ void issue_slab(MegaCtx& c, const GemmPhase& g, int tile, int k0, int Kc) {
  const uint32_t buf = (uint32_t)c.phase_issued % g.nbuf;
  const uint32_t bytes = (uint32_t)Kc * (uint32_t)(g.bn * 2);
  mbar_expect_tx(c.sm.bfull + buf, bytes);
  bulk_g2s(c.sm.a + g.slab_off + (size_t)buf * g.slab_bytes, g.w + ((size_t)tile * g.K + k0) * g.bn, bytes, c.sm.bfull + buf);
}

// issue this CTA's slabs of phase g up to item index `upto` (exclusive), in order
__device__ __forceinline__ void issue_items(MegaCtx& c, const GemmPhase& g, int upto) {
  const int items = phase_items(g);
  const int ntiles = items / g.ksplits;
  const int Kc = g.K / g.ksplits;
  while (c.phase_issued < upto) {
    const int w = blockIdx.x + c.phase_issued * gridDim.x;
    if (w >= items) break;
    issue_slab(c, g, w % ntiles, (w / ntiles) * Kc, Kc);
    ++c.phase_issued;
  }
}

// called before the barrier that precedes phase g: weights do not depend on anything, start fetching them now
__device__ __forceinline__ void prefetch_phase0(MegaCtx& c, const GemmPhase& g) {
  c.phase_consumed = 0;                   // every issuing thread keeps its own copy of the consumer-side counters
  if (threadIdx.x != 0) return;
  c.phase_issued = 0;
  issue_items(c, g, (int)g.nbuf);         // every buffer is free here: the previous GEMM phase has retired
}

template <bool PROF>
__device__ void gemm_phase0(const MegaParams& p, MegaCtx& c, const GemmPhase& g) {
  const int items = phase_items(g);
  const int ntiles = items / g.ksplits;
  const int Kc = g.K / g.ksplits;
  const int nkb = Kc / 64;
  const int a_rows = p.B <= 64 ? 64 : 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B <= 64 with mma_m64: M = 64 UMMA -- the instruction fetches only the 64 real activation rows from shared memory.
  // Measured neutral (profiles/r01/mega_phase_breakdown_v10_operand_paths.json): an N = 16 MMA costs ~80 cycles of issue
  // whatever its M, i.e. this phase is bound by the NUMBER of tcgen05.mma instructions -- which is why gemm_mode 1 below
  // turns the product around.  TMEM rows of an M = 64 accumulator: row r -> lane 32*(r/16) + r%16.
  const bool m64 = a_rows == 64;
  const uint32_t IDESC = m64 ? umma_idesc(1, 64, g.bn) : umma_idesc(1, 128, g.bn);
  int loaded_split = -1;
  int it = 0;
  float rstd = 1.0f;          // fused norms: 1 / rms of this thread's row of the activation slab
  const long long gt_entry = (PROF && p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;
  for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
    const int tile = w % ntiles, split = w / ntiles;
    // optional fine-grained timing (CTA 0, thread 0): prof[14..17] = activation slab load, weight slab wait, MMA issue,
    // commit -> end of epilogue
    const bool gprof = PROF && p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long gt = gprof ? clock64() : 0;
#define GEMM_MARK(slot_) do { if (gprof) { const long long t_ = clock64(); p.prof[slot_] += t_ - gt; gt = t_; } } while (0)
    bool a_pending = false;
    if (split != loaded_split) {          // (re)load the activation slab for this K range
      if (threadIdx.x == 0) {             // one bulk copy: k-blocks [split*Kc/64, +nkb) of the swizzled image are contiguous
#ifdef IVG_MEGA_FENCE_ALL
        asm volatile("fence.proxy.async;\n" ::: "memory");
#endif
        const uint32_t bytes = (uint32_t)(nkb * a_rows * 128);
        mbar_expect_tx(c.sm.abar, bytes);
        bulk_g2s(c.sm.a, g.A + (size_t)((split * Kc) >> 6) * (size_t)(a_rows * 64), bytes, c.sm.abar);
      }
      a_pending = true;
      loaded_split = split;
    }
    GEMM_MARK(14);
    if (threadIdx.x == 0) issue_items(c, g, it + (int)g.nbuf);      // this item (if not prefetched) and the following nbuf - 1
    if (lane == 0 && warp < MEGA_NISSUE) {
      // every issuing thread walks the same sequence of operand barriers (own copies of the parities)
      const uint32_t buf = (uint32_t)c.phase_consumed % g.nbuf;
      const uint32_t par = (c.wait_par >> buf) & 1u;
      c.wait_par ^= 1u << buf;
      ++c.phase_consumed;
      if (a_pending) mbar_wait_bounded(p, c.sm.abar, c.aphase, 3);
      mbar_wait_bounded(p, c.sm.bfull + buf, par, 4);
      tc_fence_after();
      GEMM_MARK(15);
      const uint32_t a0 = smem_u32(c.sm.a), b0 = smem_u32(c.sm.a + g.slab_off + (size_t)buf * g.slab_bytes);
      // k-step t = 4 j + k goes to issuer / accumulator t % MEGA_NISSUE (64 TMEM columns apart); the epilogue adds the
      // accumulators in a fixed order.  Each issuer commits separately: mma_done counts MEGA_NISSUE arrivals.
      const uint32_t acc = c.tmem_base + (uint32_t)(warp * 64);
      for (int j = 0; j < nkb; ++j) {
        const uint64_t adesc = umma_desc_sw128_kmajor(a0 + (uint32_t)(j * a_rows * 128));
        const uint64_t bdesc = umma_desc_sw128_kmajor(b0 + (uint32_t)(j * g.bn * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((k % MEGA_NISSUE) == warp)
            umma_ss<false>(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (j * 4 + k) >= MEGA_NISSUE ? 1u : 0u);
      }
      umma_commit(c.sm.mma_done);
      GEMM_MARK(16);
    }
    // ---- epilogue: warps 4..7 own TMEM lane quadrants 0..3 ----
    if (warp >= 4 && warp < 8) {
      const int q = warp & 3;
      const int row = m64 ? (lane < 16 ? q * 16 + lane : p.B) : q * 32 + lane;
      if (MEGA_FUSE_NORM && g.norm_a && a_pending) {
        // rstd of this thread's row from the activation slab (bf16(x), K = hidden) while the tensor core works; with M = 64
        // tiles only 16 lanes of a warp own rows, the other 16 take every second k-block of the same rows
        mbar_wait_bounded(p, c.sm.abar, c.aphase, 3);
        const int r = m64 ? q * 16 + (lane & 15) : q * 32 + lane;
        float ssq = 0.f;
        if (r < p.B) {
          for (int j = m64 ? (lane >> 4) : 0; j < nkb; j += m64 ? 2 : 1) {
            const uint4* rp = reinterpret_cast<const uint4*>(c.sm.a + (size_t)j * (size_t)(a_rows * 128) + (size_t)r * 128);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 u = rp[ch ^ (r & 7)];      // rows of a quarter-warp hit 8 different 16-byte columns: no bank conflicts
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); ssq = fmaf(f.x, f.x, ssq); ssq = fmaf(f.y, f.y, ssq); }
            }
          }
        }
        if (m64) ssq += __shfl_xor_sync(0xffffffffu, ssq, 16);
        rstd = rsqrtf(ssq / (float)g.K + p.eps);
      }
      mbar_wait_bounded(p, c.sm.mma_done, c.mphase, 5);
      tc_fence_after();
      for (int c0 = 0; c0 < g.bn; c0 += 16) {          // 16 accumulator columns (weight rows) at a time
        uint32_t r[MEGA_NACC][16];                          // all accumulators in flight, one wait
#pragma unroll
        for (int a = 0; a < MEGA_NACC; ++a)                // K >= 64 per item: every accumulator has been written
          tmem_ld_32x32b_x16(c.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + c0), r[a]);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[0][i]);
#pragma unroll
        for (int a = 1; a < MEGA_NACC; ++a) {              // fixed order: reproducible
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(r[a][i]);
        }
        if (MEGA_FUSE_NORM && g.norm_a) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= rstd;
        }
        if (row < p.B) {
          const int n0 = tile * g.bn + c0;
          if (g.epi == EPI_STORE_BF16) {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)row * g.ldo + n0);
            op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
          } else if (g.epi == EPI_PARTIAL_F32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + ((size_t)split * p.B + row) * g.ldo + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else if (g.epi == EPI_SWIGLU) {
            if (n0 < g.N) {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + a_off(p, row, n0 >> 1, g.ldo));
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = silu_f(v[2 * i]) * v[2 * i + 1];
              op[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            }
          } else {
            float* op = reinterpret_cast<float*>(g.out) + (size_t)row * g.ldo + n0;
#pragma unroll
            for (int i = 0; i < 16; ++i) if (n0 + i < g.N) op[i] = v[i];
          }
        }
      }
      tc_fence_before();
    }
    c.mphase ^= 1;          // every thread tracks the parity (only warps 4..7 wait on it)
    if (a_pending) c.aphase ^= 1u;   // ... and the parity of the activation-slab barrier (issuers and epilogue warps wait on it)
    if (MEGA_FUSE_NORM && g.resid_cnt != nullptr && warp >= 4) __threadfence();   // partials of this item: released before the count
    __syncthreads();        // MMA retired (epilogue observed mma_done): TMEM accumulator, A slab and this weight
                            // buffer may be reused
    if (MEGA_FUSE_NORM && g.resid_cnt != nullptr) {
      // x += sum of the split-K partials of this output tile, by the LAST of its `ksplits` work items to arrive
      __shared__ unsigned int s_last;
      if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(g.resid_cnt + tile, 1u);
        __threadfence();
        s_last = ((old + 1u) % (unsigned int)g.ksplits == 0u) ? 1u : 0u;
      }
      __syncthreads();
      if (s_last) {
        const int f4_per_row = g.bn >> 2;                                   // float4 groups per row of the tile
        for (int e = threadIdx.x; e < p.B * f4_per_row; e += MEGA_THREADS) {
          const int rr = e / f4_per_row, cc = tile * g.bn + (e - rr * f4_per_row) * 4;
          if (cc < g.N) {
            float* xr = p.x + (size_t)rr * g.ldo + cc;
            float4 v = __ldcg(reinterpret_cast<const float4*>(xr));           // past L1: another SM wrote it a phase ago
            float4 q[8];
#pragma unroll
            for (int s = 0; s < 8; ++s)                                   // every load in flight before the first add; past L1
              if (s < g.ksplits) q[s] = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.out) + ((size_t)s * p.B + rr) * g.ldo + cc));
#pragma unroll
            for (int s = 0; s < 8; ++s)
              if (s < g.ksplits) { v.x += q[s].x; v.y += q[s].y; v.z += q[s].z; v.w += q[s].w; }   // split order: reproducible
            *reinterpret_cast<float4*>(xr) = v;
            uint2 o;
            o.x = pack_bf16x2(v.x, v.y);
            o.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(p.xn + a_off(p, rr, cc, g.ldo)) = o;  // un-normalised image: the consumer applies rstd
          }
        }
      }
      // (the device-wide barrier that follows orders these writes before the next phase; its proxy fence covers the image)
    }
    GEMM_MARK(17);
#undef GEMM_MARK
  }
  if (PROF && p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.prof[18] += clock64() - gt_entry;
}

// =====================================================================================================================
// gemm_mode 1: weight-stationary GEMM phases.
// Measured on the mode-0 phases above (profiles/r01/mega_phase_breakdown_v9/v10): a tcgen05.mma with N = 16 costs ~80
// cycles of issue on the single issuing thread whatever M is, so a step's ~3600 MMAs alone took ~165 us, and every CTA
// re-read the whole [B, K] activation matrix from L2 for each phase.  Here the product is turned around:
//     D[64 weight rows, a_rows batch columns] += W_tile[64, 16] . X[a_rows, 16]^T        (M = 64, N = a_rows, K = 16)
// one instruction now covers 64 weight rows instead of 16 (~3.4x fewer instructions per step), a work item is
// (64-row tile, K split) so a CTA loads only the K range of the activations it needs (one bulk copy of the swizzled
// image, see a_off), and weights stream through 32 KB slabs of 64 rows x <=256 k with up to nbuf in flight.
//   qkv    : 3h/64 tiles x qkv_splits        -> fp32 partials [split][B][3h], summed by the attention prologue
//   o/down : h/64 tiles x o_splits/d_splits  -> fp32 partials [split][B][h], summed by the norm phase (fixed order)
//   gate/up: 2*inter/64 tiles, whole K streamed in chunks, SwiGLU in the epilogue (weights packed so that gate_i / up_i
//            sit 8 TMEM lanes apart in the same warp: one shuffle)
//   lm_head: ceil(vocab/64) tiles, whole K streamed.
// Accumulator row r (weight row) lives in TMEM lane 32*(r/16) + r%16, column b = batch row.
// =====================================================================================================================
__device__ __forceinline__ int w64_chunk(int Ks) {        // K per weight slab: largest multiple of 64 <= MEGA_W_CHUNK dividing Ks
  return Ks % MEGA_W_CHUNK == 0 ? MEGA_W_CHUNK : 64;
}
__device__ __forceinline__ int w64_items(const GemmPhase& g) { return ((g.N + MEGA_WM - 1) / MEGA_WM) * g.ksplits; }

// issue this CTA's weight slabs of phase g up to slab index `upto` (exclusive); slab s = (item s / nch, chunk s % nch)
__device__ __forceinline__ void w64_issue(const MegaParams& p, MegaCtx& c, const GemmPhase& g, int upto) {
  const int items = w64_items(g);
  const int ntiles = items / g.ksplits;
  const int Ks = g.K / g.ksplits;
  const int chunk = w64_chunk(Ks);
  const int nch = Ks / chunk;
  while (c.phase_issued < upto) {
    const int it = c.phase_issued / nch, ch = c.phase_issued - it * nch;
    const int w = blockIdx.x + it * gridDim.x;
    if (w >= items) break;
    const int tile = w % ntiles, split = w / ntiles;
    const uint32_t buf = c.issued % c.sm.nbuf;
    const uint32_t use = c.issued / c.sm.nbuf;
    if (use > 0) mbar_wait_bounded(p, c.sm.bempty + buf, (use - 1) & 1u, 6);      // previous occupant's MMAs retired
    ++c.issued;
    const uint32_t bytes = (uint32_t)chunk * 128u;                                 // 64 rows x chunk x 2 B
    mbar_expect_tx(c.sm.bfull + buf, bytes);
    const size_t kb0 = (size_t)(split * Ks + ch * chunk) >> 6;
    bulk_g2s(c.sm.b0 + (size_t)buf * c.sm.slab_bytes, g.w + ((size_t)tile * (size_t)(g.K >> 6) + kb0) * (MEGA_WM * 64), bytes,
             c.sm.bfull + buf);
    ++c.phase_issued;
  }
}

template <bool PROF>
__device__ void gemm_phase_w64(const MegaParams& p, MegaCtx& c, const GemmPhase& g) {
  const int items = w64_items(g);
  const int ntiles = items / g.ksplits;
  const int Ks = g.K / g.ksplits;
  const int chunk = w64_chunk(Ks);
  const int nch = Ks / chunk;
  const int a_rows = p.a_rows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t IDESC = umma_idesc(1, MEGA_WM, a_rows);
  int loaded_split = -1;
  int it = 0;
  // optional timing (CTA 0, thread 0), accumulated in registers and flushed once per phase: prof[14..17] = activation
  // copy issue, slab wait, MMA issue, commit -> end of epilogue; prof[18] = whole phase, entry to exit
  const bool gprof = PROF && p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long gacc0 = 0, gacc1 = 0, gacc2 = 0, gacc3 = 0;
  const long long gt_entry = gprof ? clock64() : 0;
  for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
    const int tile = w % ntiles, split = w / ntiles;
    long long gt = gprof ? clock64() : 0;
#define GEMM_MARK(acc_) do { if (gprof) { const long long t_ = clock64(); acc_ += t_ - gt; gt = t_; } } while (0)
    if (threadIdx.x == 0) {
      bool a_pending = false;
      if (split != loaded_split) {        // this split's K range of the activation image: contiguous k-blocks, one bulk copy
#ifdef IVG_MEGA_FENCE_ALL
        asm volatile("fence.proxy.async;\n" ::: "memory");
#endif
        const uint32_t bytes = (uint32_t)((Ks >> 6) * a_rows * 128);
        mbar_expect_tx(c.sm.abar, bytes);
        bulk_g2s(c.sm.a, g.A + (size_t)((split * Ks) >> 6) * (size_t)(a_rows * 64), bytes, c.sm.abar);
        a_pending = true;
      }
      GEMM_MARK(gacc0);
      const uint32_t a0 = smem_u32(c.sm.a);
      for (int ch = 0; ch < nch; ++ch) {
        w64_issue(p, c, g, it * nch + ch + (int)c.sm.nbuf);     // this slab (if not prefetched) and the next nbuf - 1
        const uint32_t buf = c.consumed % c.sm.nbuf;
        const uint32_t par = (c.consumed / c.sm.nbuf) & 1u;
        ++c.consumed;
        if (a_pending) { mbar_wait_bounded(p, c.sm.abar, c.aphase, 3); c.aphase ^= 1u; a_pending = false; }
        mbar_wait_bounded(p, c.sm.bfull + buf, par, 4);
        tc_fence_after();
        GEMM_MARK(gacc1);
        const uint32_t w0 = smem_u32(c.sm.b0 + (size_t)buf * c.sm.slab_bytes);
        const int nkb = chunk >> 6;
        for (int j = 0; j < nkb; ++j) {
          const uint64_t wdesc = umma_desc_sw128_kmajor(w0 + (uint32_t)(j * MEGA_WM * 128));
          const uint64_t xdesc = umma_desc_sw128_kmajor(a0 + (uint32_t)((ch * nkb + j) * a_rows * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss<false>(c.tmem_base, wdesc + (uint64_t)(k * 2), xdesc + (uint64_t)(k * 2), IDESC, (ch | j | k) ? 1u : 0u);
        }
        umma_commit(c.sm.bempty + buf);                         // slab buffer reusable once these MMAs retire
        GEMM_MARK(gacc2);
      }
      umma_commit(c.sm.mma_done);
    }
    loaded_split = split;
    // ---- epilogue: warps 4..7 own TMEM lane quadrants 0..3; lanes 0..15 of a quadrant hold weight rows 16 q + lane ----
    if (warp >= 4 && warp < 8) {
      const int q = warp & 3;
      const int R = q * 16 + (lane & 15);
      const bool lane_ok = lane < 16;
      const int n = tile * MEGA_WM + R;
      mbar_wait_bounded(p, c.sm.mma_done, c.mphase, 5);
      tc_fence_after();
      for (int c0 = 0; c0 < a_rows; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(c.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (g.epi == EPI_SWIGLU) {
          // packed row order inside each 16-row group: 8 gate rows then the 8 matching up rows (mega_pack_weight64)
          const int i = tile * (MEGA_WM / 2) + q * 8 + (lane & 7);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float gv = __uint_as_float(r[j]);
            const float uv = __shfl_down_sync(0xffffffffu, gv, 8);
            const int b = c0 + j;
            if (lane < 8 && b < p.B && i < (g.N >> 1))
              reinterpret_cast<__nv_bfloat16*>(g.out)[a_off(p, b, i, g.ldo)] = __float2bfloat16_rn(silu_f(gv) * uv);
          }
        } else {
          float* op = reinterpret_cast<float*>(g.out) + (g.epi == EPI_PARTIAL_F32 ? (size_t)split * p.B * g.ldo : (size_t)0) + n;
          if (lane_ok && n < g.N) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.B) op[(size_t)(c0 + j) * g.ldo] = __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
    }
    c.mphase ^= 1;
    __syncthreads();        // accumulator drained, activation slab reusable
    GEMM_MARK(gacc3);
#undef GEMM_MARK
  }
  if (gprof) {
    p.prof[14] += gacc0; p.prof[15] += gacc1; p.prof[16] += gacc2; p.prof[17] += gacc3;
    p.prof[18] += clock64() - gt_entry;
  }
}

// GM (GEMM mode) and AM (attention mode) are template parameters of the kernel: one instantiation holds exactly one GEMM
// path and one attention path.  (With both paths behind runtime branches the 255-register kernel spilled and the hot
// loops of the paths that were NOT taken got slower: 238 -> 286 ms per rollout, profiles/r01 v10 -> v11.)
template <int GM>
__device__ __forceinline__ void prefetch_phase(const MegaParams& p, MegaCtx& c, const GemmPhase& g) {
  if constexpr (GM == 0) { prefetch_phase0(c, g); return; }
  if (threadIdx.x != 0) return;
  c.phase_issued = 0;
  w64_issue(p, c, g, (int)c.sm.nbuf);
}
template <int GM, bool PROF>
__device__ __forceinline__ void gemm_phase(const MegaParams& p, MegaCtx& c, const GemmPhase& g) {
  if constexpr (GM == 0) gemm_phase0<PROF>(p, c, g); else gemm_phase_w64<PROF>(p, c, g);
}

// ---- add split-K partials (fixed order) + RMSNorm -> xn ; or embedding gather + RMSNorm ----
// One row per CTA, one float4 per thread (hidden <= 1024): every load of the row is in flight at once.
// raw (fused norms): only x = E[token] (+ slot embedding) and its bf16 image; the consumer applies rstd, g is folded into its weights
template <int MAXP>
__device__ void norm_phase(const MegaParams& p, const float* w, int nparts, const long long* tok_row0, int tok_col, bool raw = false) {
  __shared__ float s_ss[MEGA_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.hidden;
  const int i = threadIdx.x * 4;
  for (int m = blockIdx.x; m < p.B; m += gridDim.x) {
    float* xr = p.x + (size_t)m * H;
    const float* src = xr;
    if (tok_row0) {
      long long id = tok_row0[(size_t)m * p.tok_stride + tok_col];
      id = id < 0 ? 0 : (id >= p.vocab ? p.vocab - 1 : id);
      src = p.embed + (size_t)id * H;
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float ss = 0.f;
    if (i < H) {
      v = *reinterpret_cast<const float4*>(src + i);
#ifndef IVG_MEGA_NO_SLOTS
      if (tok_row0 && p.slot_emb && p.slot_period > 0 && tok_col >= p.slot0 && (tok_col - p.slot0) % p.slot_period == 0 &&
          (tok_col - p.slot0) / p.slot_period < p.nslots) {       // forced slot: + action_linear(a_i), action_model.py:80-81
        const float4 e = *reinterpret_cast<const float4*>(
            p.slot_emb + ((size_t)m * p.nslots + (tok_col - p.slot0) / p.slot_period) * H + i);
        v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
      }
#endif
      float4 q[MAXP];
#pragma unroll
      for (int s = 0; s < MAXP; ++s)
        if (s < nparts) q[s] = *reinterpret_cast<const float4*>(p.part + ((size_t)s * p.B + m) * H + i);
#pragma unroll
      for (int s = 0; s < MAXP; ++s)
        if (s < nparts) { v.x += q[s].x; v.y += q[s].y; v.z += q[s].z; v.w += q[s].w; }   // fixed order: reproducible
      *reinterpret_cast<float4*>(xr + i) = v;
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    if (raw) {                 // CTA-uniform
      if (i < H) {
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(p.xn + a_off(p, m, i, H)) = o;
      }
      continue;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if (lane == 0) s_ss[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < MEGA_THREADS / 32; ++k) tot += s_ss[k];
    const float r = rsqrtf(tot / (float)H + p.eps);
    if (i < H) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(w + i));
      uint2 o;
      o.x = pack_bf16x2(g.x * (v.x * r), g.y * (v.y * r));
      o.y = pack_bf16x2(g.z * (v.z * r), g.w * (v.w * r));
      *reinterpret_cast<uint2*>(p.xn + a_off(p, m, i, H)) = o;
    }
    __syncthreads();
  }
}

// q/k/v input of the attention prologue: the bf16 row of mode 0, or the sum of the qkv projection's split-K partials
// (fixed order) rounded to bf16 exactly like the stored row would have been
constexpr int MEGA_MAX_QKV_SPLITS = 4;
template <int GM>
__device__ __forceinline__ float qkv_in(const MegaParams& p, int b, int col) {
  if constexpr (GM == 0) return __bfloat162float(p.qkv[(size_t)b * 3 * p.hidden + col]);
  const size_t ld = (size_t)3 * p.hidden;
  float part[MEGA_MAX_QKV_SPLITS];
#pragma unroll
  for (int sp = 0; sp < MEGA_MAX_QKV_SPLITS; ++sp)          // every load in flight before the first add
    part[sp] = sp < p.qkv_splits ? p.qkvp[((size_t)sp * p.B + b) * ld + col] : 0.f;
  float s = part[0];
#pragma unroll
  for (int sp = 1; sp < MEGA_MAX_QKV_SPLITS; ++sp) s += part[sp];
  return __bfloat162float(__float2bfloat16_rn(s));
}

// ---- RoPE + KV append + attention over the cache: ONE WARP per (b, head) item (used with 256-thread CTAs) ----
// Measured history of this phase (B=64, 12 heads, L 514..750, us per layer on CTA 0): CTA per item 59; warp per item
// with 16 register-staged 16-byte loads per lane 51-53 (this version); warp pairs / 512 threads 55; L2-only loads 53;
// cp.async ring with a badly banked layout 148; splitting items in thirds over all warps cannot help (the busiest warp
// still runs ~22 dependent load batches).  At ~2.4 us per dependent batch the phase streams 2.3 TB/s; the standalone
// fused kernel (768 CTAs, no unrolling) lands on the same 50-62 us.  Closing the gap to HBM speed needs ~100 KB of
// loads in flight per SM, i.e. TMA bulk copies into a large shared-memory ring -- next round.
constexpr int ATW_QB = 16;    // K rows in flight per lane group (8 lanes per row, 4 rows per pass)
constexpr int ATW_DB = 16;    // V^T rows in flight per lane
template <int GM>
__device__ void attention_warp(const MegaParams& p, int layer, int bh, int pos, float* wsm) {
  const int heads = p.heads, Lmax = p.Lmax, Hd = p.hidden;
  const int Lcur = pos + 1;
  float* sc = wsm;                    // [Lmax + 8]
  float* qs = wsm + Lmax + 8;         // [64]
  const int b = bh / heads, hh = bh - b * heads;
  const int lane = threadIdx.x & 31;
  __nv_bfloat16* kslab = p.kcache + ((size_t)layer * p.B * heads + bh) * Lmax * 64;
  __nv_bfloat16* vslab = p.vcache + ((size_t)layer * p.B * heads + bh) * 64 * Lmax;
  {
    const float cs = __ldg(p.cos_tab + (size_t)pos * 32 + lane), sn = __ldg(p.sin_tab + (size_t)pos * 32 + lane);
    const float q0 = qkv_in<GM>(p, b, hh * 64 + lane), q1 = qkv_in<GM>(p, b, hh * 64 + lane + 32);
    const float k0 = qkv_in<GM>(p, b, Hd + hh * 64 + lane), k1 = qkv_in<GM>(p, b, Hd + hh * 64 + lane + 32);
    const __nv_bfloat16 v0 = __float2bfloat16_rn(qkv_in<GM>(p, b, 2 * Hd + hh * 64 + lane));
    const __nv_bfloat16 v1 = __float2bfloat16_rn(qkv_in<GM>(p, b, 2 * Hd + hh * 64 + lane + 32));
    qs[lane] = __bfloat162float(__float2bfloat16_rn(q0 * cs - q1 * sn)) * 0.125f;
    qs[lane + 32] = __bfloat162float(__float2bfloat16_rn(q1 * cs + q0 * sn)) * 0.125f;
    kslab[(size_t)pos * 64 + lane] = __float2bfloat16_rn(k0 * cs - k1 * sn);
    kslab[(size_t)pos * 64 + lane + 32] = __float2bfloat16_rn(k1 * cs + k0 * sn);
    vslab[(size_t)lane * Lmax + pos] = v0;
    vslab[(size_t)(lane + 32) * Lmax + pos] = v1;
  }
  __syncwarp();
  const int sub = lane & 7, rslot = lane >> 3;
  float qreg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qreg[i] = qs[sub * 8 + i];
  float mx = -INFINITY;
  for (int l0 = 0; l0 < Lcur; l0 += 4 * ATW_QB) {
    uint4 kv[ATW_QB];
#pragma unroll
    for (int u = 0; u < ATW_QB; ++u) {
      const int l = l0 + u * 4 + rslot;
      kv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (l < Lcur) kv[u] = ldg_cg_v4(kslab + (size_t)l * 64 + sub * 8);
    }
#pragma unroll
    for (int u = 0; u < ATW_QB; ++u) {
      const int l = l0 + u * 4 + rslot;
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[u]);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        part = fmaf(qreg[2 * i], f.x, part);
        part = fmaf(qreg[2 * i + 1], f.y, part);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      if (l < Lcur) {
        mx = fmaxf(mx, part);
        if (sub == 0) sc[l] = part;
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  __syncwarp();
  float sum = 0.f;
  for (int l = lane; l < Lcur + 8; l += 32) {
    const float e = l < Lcur ? __expf(sc[l] - mx) : 0.f;
    sc[l] = e;
    sum += e;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float inv = 1.0f / sum;
  __syncwarp();
  for (int d0 = 0; d0 < 64; d0 += ATW_DB) {
    float acc[ATW_DB];
#pragma unroll
    for (int r = 0; r < ATW_DB; ++r) acc[r] = 0.f;
    for (int l = lane * 8; l < Lcur; l += 256) {
      uint4 vv[ATW_DB];
#pragma unroll
      for (int r = 0; r < ATW_DB; ++r) vv[r] = ldg_cg_v4(vslab + (size_t)(d0 + r) * Lmax + l);
      float pr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pr[i] = sc[l + i];
      const int nvalid = Lcur - l;
#pragma unroll
      for (int r = 0; r < ATW_DB; ++r) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vv[r]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h2[i]);
          acc[r] += (2 * i < nvalid) ? pr[2 * i] * f.x : 0.f;
          acc[r] += (2 * i + 1 < nvalid) ? pr[2 * i + 1] * f.y : 0.f;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ATW_DB; ++r) {
      float a = acc[r];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      if (lane == r) p.ao[a_off(p, b, hh * 64 + d0 + r, Hd)] = __float2bfloat16_rn(a * inv);
    }
  }
  __syncwarp();
}

// ring-slot wait that can never hang the GPU: after ~1 s the error flag is raised (the host raises) and bit 31 of
// `par` makes every later wait of this warp fall through.
__device__ __forceinline__ void ring_wait(const MegaParams& p, uint64_t* bar, uint32_t& par, int slot) {
  if (!(par >> 31)) {
    const uint32_t phase = (par >> slot) & 1u;
    if (!mbar_try_wait(bar, phase)) {
      const long long t0 = clock64();
      while (!mbar_try_wait(bar, phase)) {
        if (clock64() - t0 > (1ll << 31)) { *p.error = 2; par |= 0x80000000u; break; }
      }
    }
    if (__any_sync(0xffffffffu, (par >> 31) != 0u)) par |= 0x80000000u;
  }
  par ^= 1u << slot;
}

// ---- RoPE + KV append + attention over the cache: one warp per (b, head) item, K/V streamed through a ring ----
// The register-staged version above stalls once per batch of loads (~22 dependent batches of ~2.4 us per item).  Here
// lane 0 keeps `nslot` 4 KB bulk copies in flight per warp (the activation region is split between the warps that own
// work this phase), the copy engine writes shared memory directly and the warp consumes slot after slot.  Both K and V
// are read from [Lmax][64] slabs (V from the row-major second cache `vrows`), 32 positions = 4 KB per unit:
//   K units : rows -> scores (8 lanes per row, butterfly over the 8 lanes; same arithmetic as above)
//   V units : lane (sub, rslot) accumulates dims [8 sub, 8 sub + 8) over positions = rslot mod 4; no cross-lane
//             traffic until one 2-step reduction at the end.
// V units do not depend on the softmax, so they are already in flight while the scores are being computed.  The row
// of the token being decoded never goes through the ring: its score and value come from registers, so the first
// copies (old rows only) are issued before the RoPE arithmetic.
// A call handles K/V units [u0, u1) of the item and, when `tail` is set, the new row (and the K/V append); it returns
// the flash-decoding partial (max, sum, unnormalised accumulator) -- whole items use [0, nK) with tail, the items
// left over after an even deal (768 items on 148 SMs leave 28) are cut into MEGA_ATT_SPLIT parts on otherwise idle warps.
// (History: V^T rows + a butterfly per row 61 us/layer; this layout 34.7; 5.5 TB/s is what the ring sustains in
// isolation, tools/probes/ring_probe.cu.)
constexpr int MEGA_ATT_SPLIT = 4;
constexpr int MEGA_ATT_MAXSPLIT = 8;
// Small batches (B * heads * 2 <= 8 * #CTAs, e.g. cfg256's B = 16: 192 items on 148 CTAs): one warp per item leaves most warps
// idle and a single warp keeps only ~32 KB of K/V in flight (latency-bound: 22 us per layer for 37 MB).  Then EVERY item is
// cut into P = floor(8 * #CTAs / items) <= 8 parts along the sequence, dealt round-robin over all warps of all CTAs, and
// merged by the last part to arrive exactly like the left-over items of the large-batch deal.
__device__ __forceinline__ int att_uniform_parts(const MegaParams& p) {
  const int total = p.B * p.heads, G = (int)gridDim.x;
  int P = (8 * G) / total;
  P = P > MEGA_ATT_MAXSPLIT ? MEGA_ATT_MAXSPLIT : P;
  return P >= 2 ? P : 0;
}

// unit u of the item's [u0, u0 + nU) K units followed by the same V units -> ring slot (lane 0 only; rows < pos were
// written in earlier steps)
__device__ __forceinline__ void att_issue(const MegaParams& p, int layer, int bh, int pos, int u0, int nU, int u, int slot,
                                          uint8_t* ring, uint64_t* bars) {
  const size_t slab = ((size_t)layer * p.B * p.heads + bh) * p.Lmax * 64;
  const int c = u0 + (u < nU ? u : u - nU);
  const int r0 = c << 5;
  const uint32_t bytes = (uint32_t)(pos - r0 < 32 ? pos - r0 : 32) * 128u;
  mbar_expect_tx(bars + slot, bytes);
  bulk_g2s(ring + (size_t)slot * MEGA_RING_SLOT, (u < nU ? p.kcache : p.vrows) + slab + (size_t)r0 * 64, bytes, bars + slot);
}

// what a warp does in the attention phase of this step
struct AttWork {
  int kind;            // 0 nothing, 1 whole item, 2 part of an item cut along the sequence
  int bh, u0, u1, extra, q, nslot;
  int nsplit;          // parts per cut item (MEGA_ATT_SPLIT for the left-over items of the large-batch deal)
  bool tail;
};

// even deal: `full` whole items per CTA, the left-over items cut in MEGA_ATT_SPLIT parts on otherwise idle warps
__device__ __forceinline__ bool att_even_deal(const MegaParams& p) {
  const int total = p.B * p.heads, G = (int)gridDim.x;
  const int full = total / G, extras = total - full * G, free_w = 8 - full;
  if (att_uniform_parts(p) != 0) return true;
  return full <= 8 && (extras == 0 || (free_w > 0 && extras * MEGA_ATT_SPLIT <= free_w * G));
}
__device__ __forceinline__ AttWork att_assign(const MegaParams& p, uint32_t a_bytes, int warp, int pos) {
  const int total = p.B * p.heads, G = (int)gridDim.x;
  const int P = att_uniform_parts(p);
  if (P != 0) {                       // every item in P parts, part pi = item * P + q on CTA pi % G, warp pi / G
    AttWork w;
    w.kind = 0;
    const int nparts_u = total * P;
    const int mine = nparts_u > (int)blockIdx.x ? (nparts_u - (int)blockIdx.x + G - 1) / G : 0;     // <= 8 by construction
    w.nslot = mine > 0 ? (int)(a_bytes / MEGA_RING_SLOT) / mine : 1;
    w.nslot = w.nslot > 8 ? 8 : w.nslot;
    w.nsplit = P;
    const int nK = (pos + 31) >> 5;
    if (warp < mine) {
      const int pi = (int)blockIdx.x + G * warp;
      w.kind = 2; w.bh = pi / P; w.q = pi - w.bh * P; w.extra = w.bh;
      w.u0 = (nK * w.q) / P; w.u1 = (nK * (w.q + 1)) / P; w.tail = w.q == P - 1;
    }
    return w;
  }
  const int full = total / G, extras = total - full * G;
  const int nparts = extras * MEGA_ATT_SPLIT;
  const int my_parts = nparts > (int)blockIdx.x ? (nparts - (int)blockIdx.x + G - 1) / G : 0;
  const int active = full + my_parts;
  AttWork w;
  w.kind = 0;
  w.nsplit = MEGA_ATT_SPLIT;
  w.nslot = active > 0 ? (int)(a_bytes / MEGA_RING_SLOT) / active : 1;     // o-proj slabs sit above a_bytes
  w.nslot = w.nslot > 8 ? 8 : w.nslot;
  const int nK = (pos + 31) >> 5;
  if (warp < full) {
    w.kind = 1; w.bh = (int)blockIdx.x + G * warp; w.u0 = 0; w.u1 = nK; w.tail = true; w.extra = 0; w.q = 0;
  } else if (warp < active) {
    const int pi = (int)blockIdx.x + G * (warp - full);
    w.kind = 2; w.extra = pi / MEGA_ATT_SPLIT; w.q = pi % MEGA_ATT_SPLIT; w.bh = full * G + w.extra;
    w.u0 = (nK * w.q) / MEGA_ATT_SPLIT; w.u1 = (nK * (w.q + 1)) / MEGA_ATT_SPLIT; w.tail = w.q == MEGA_ATT_SPLIT - 1;
  }
  return w;
}
// first ring fill: old K/V rows do not depend on this step's qkv projection, so it is issued BEFORE the device-wide
// barrier that precedes the attention phase and lands while the CTA waits there
__device__ __forceinline__ void attention_prefetch(const MegaParams& p, int layer, int pos, const AttWork& w, uint8_t* ring,
                                                   uint64_t* bars) {
  if (w.kind == 0 || (threadIdx.x & 31) != 0) return;
  const int nU = w.u1 - w.u0, units = 2 * nU;
  const int pre = units < w.nslot ? units : w.nslot;
  for (int u = 0; u < pre; ++u) att_issue(p, layer, w.bh, pos, w.u0, nU, u, u, ring, bars);
}

template <int GM, bool PROF>
__device__ void attention_stream(const MegaParams& p, int layer, int bh, int pos, int u0, int u1, bool tail, bool issued,
                                 float* wsm, uint8_t* ring, int nslot, uint64_t* bars, uint32_t& par, float& m_out,
                                 float& l_out, float (&acc)[8]) {
  const int heads = p.heads, Lmax = p.Lmax, Hd = p.hidden;
  float* sc = wsm;                    // [Lmax + 8] (indexed from row 32 * u0)
  float* qs = wsm + Lmax + 8;         // [64] q, later the new row's v
  const int b = bh / heads, hh = bh - b * heads;
  const int lane = threadIdx.x & 31;
  __nv_bfloat16* kslab = p.kcache + ((size_t)layer * p.B * heads + bh) * Lmax * 64;
  __nv_bfloat16* vslab = p.vrows + ((size_t)layer * p.B * heads + bh) * Lmax * 64;
  const int nU = u1 - u0;
  const int units = 2 * nU;
  const int row0 = u0 << 5;
  // optional fine-grained timing of the phase (CTA 0, warp 0, lane 0): prof[9..13] = prologue, K loop, softmax, V loop, tail
  const bool tprof = PROF && p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long tm = tprof ? clock64() : 0;
#define ATT_MARK(slot_) do { if (tprof) { const long long t_ = clock64(); p.prof[slot_] += t_ - tm; tm = clock64(); } } while (0)
  auto issue = [&](int u, int slot) { att_issue(p, layer, bh, pos, u0, nU, u, slot, ring, bars); };
  if (lane == 0 && !issued) {
    const int pre = units < nslot ? units : nslot;
    for (int u = 0; u < pre; ++u) issue(u, u);
  }
  float s_new = -INFINITY;
  float v0f = 0.f, v1f = 0.f;
  __nv_bfloat16 ka, kb, v0, v1;       // the new K/V row (this lane's two dims); appended to the caches at the very end
  {
    const float cs = __ldg(p.cos_tab + (size_t)pos * 32 + lane), sn = __ldg(p.sin_tab + (size_t)pos * 32 + lane);
    const float q0 = qkv_in<GM>(p, b, hh * 64 + lane), q1 = qkv_in<GM>(p, b, hh * 64 + lane + 32);
    const float k0 = qkv_in<GM>(p, b, Hd + hh * 64 + lane), k1 = qkv_in<GM>(p, b, Hd + hh * 64 + lane + 32);
    v0 = __float2bfloat16_rn(qkv_in<GM>(p, b, 2 * Hd + hh * 64 + lane));
    v1 = __float2bfloat16_rn(qkv_in<GM>(p, b, 2 * Hd + hh * 64 + lane + 32));
    const float qa = __bfloat162float(__float2bfloat16_rn(q0 * cs - q1 * sn)) * 0.125f;
    const float qb = __bfloat162float(__float2bfloat16_rn(q1 * cs + q0 * sn)) * 0.125f;
    qs[lane] = qa;
    qs[lane + 32] = qb;
    ka = __float2bfloat16_rn(k0 * cs - k1 * sn); kb = __float2bfloat16_rn(k1 * cs + k0 * sn);
    if (tail) {
      v0f = __bfloat162float(v0); v1f = __bfloat162float(v1);
      float sn_ = fmaf(qa, __bfloat162float(ka), qb * __bfloat162float(kb));
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) sn_ += __shfl_xor_sync(0xffffffffu, sn_, off);
      s_new = sn_;
    }
  }
  __syncwarp();
  const int sub = lane & 7, rslot = lane >> 3;
  float qreg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qreg[i] = qs[sub * 8 + i];
  __syncwarp();
  if (tail) { qs[lane] = v0f; qs[lane + 32] = v1f; }          // q is in registers now: the slot carries the new v row
  float mx = s_new;
  int slot = 0;
  ATT_MARK(9);
  for (int u = 0; u < nU; ++u) {
    ring_wait(p, bars + slot, par, slot);
    const uint8_t* src = ring + (size_t)slot * MEGA_RING_SLOT + rslot * 128 + sub * 16;
    uint4 kv[8];
#pragma unroll
    for (int ps = 0; ps < 8; ++ps) kv[ps] = *reinterpret_cast<const uint4*>(src + ps * 512);
    __syncwarp();                                             // every lane has read the slot: refill it
    if (lane == 0 && u + nslot < units) issue(u + nslot, slot);
    float part[8];
#pragma unroll
    for (int ps = 0; ps < 8; ++ps) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[ps]);
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        a = fmaf(qreg[2 * i], f.x, a);
        a = fmaf(qreg[2 * i + 1], f.y, a);
      }
      part[ps] = a;
    }
    // 8 rows x 8 lanes -> transposed butterfly (7 shuffles instead of 24): afterwards lane `sub` holds the full dot
    // product of pass ps == sub.  Same pairing order (4, 2, 1) as the plain butterfly, so the sums are bit-identical.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = (sub & 4) ? part[i] : part[i + 4];
      const float keep = (sub & 4) ? part[i + 4] : part[i];
      part[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = (sub & 2) ? part[i] : part[i + 2];
      const float keep = (sub & 2) ? part[i + 2] : part[i];
      part[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
      const float send = (sub & 1) ? part[0] : part[1];
      const float keep = (sub & 1) ? part[1] : part[0];
      part[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    const int ll = (u << 5) + sub * 4 + rslot;                // row index relative to row0
    if (row0 + ll < pos) {
      mx = fmaxf(mx, part[0]);
      sc[ll] = part[0];
    }
    slot = slot + 1 == nslot ? 0 : slot + 1;
  }
  ATT_MARK(10);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  __syncwarp();
  float sum = 0.f;
#pragma unroll 4
  for (int ll = lane; ll < (nU << 5); ll += 32) {             // zero-fills the tail of the last 32-row unit
    const float e = (row0 + ll < pos) ? __expf(sc[ll] - mx) : 0.f;
    sc[ll] = e;
    sum += e;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float e_new = tail ? __expf(s_new - mx) : 0.f;
  sum += e_new;
  __syncwarp();
  ATT_MARK(11);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int u = 0; u < nU; ++u) {
    ring_wait(p, bars + slot, par, slot);
    const uint8_t* src = ring + (size_t)slot * MEGA_RING_SLOT + rslot * 128 + sub * 16;
    const int l0 = (u << 5) + rslot;
    uint4 vv[8];
    float pl[8];
#pragma unroll
    for (int ps = 0; ps < 8; ++ps) {
      vv[ps] = *reinterpret_cast<const uint4*>(src + ps * 512);
      pl[ps] = sc[l0 + ps * 4];
    }
    __syncwarp();
    if (lane == 0 && nU + u + nslot < units) issue(nU + u + nslot, slot);
#pragma unroll
    for (int ps = 0; ps < 8; ++ps) {
      if (row0 + l0 + ps * 4 < pos) {                         // rows past the end of the last unit hold stale bytes
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vv[ps]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h2[i]);
          acc[2 * i] = fmaf(pl[ps], f.x, acc[2 * i]);
          acc[2 * i + 1] = fmaf(pl[ps], f.y, acc[2 * i + 1]);
        }
      }
    }
    slot = slot + 1 == nslot ? 0 : slot + 1;
  }
  ATT_MARK(12);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (tail) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(e_new, qs[sub * 8 + i], acc[i]);
  }
  m_out = mx;
  l_out = sum;
  if (tail) {                         // append the new row; later steps read it through the copy engine (async proxy)
    kslab[(size_t)pos * 64 + lane] = ka;
    kslab[(size_t)pos * 64 + lane + 32] = kb;
    vslab[(size_t)pos * 64 + lane] = v0;
    vslab[(size_t)pos * 64 + lane + 32] = v1;
    fence_proxy_async_all();
  }
  __syncwarp();
  ATT_MARK(13);
#undef ATT_MARK
}

__device__ __forceinline__ void attention_store(const MegaParams& p, int bh, const float (&acc)[8], float inv) {
  const int b = bh / p.heads, hh = bh - b * p.heads;
  const int lane = threadIdx.x & 31;
  if ((lane >> 3) == 0) {
    const uint4 o = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                               pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
    *reinterpret_cast<uint4*>(p.ao + a_off(p, b, hh * 64 + (lane & 7) * 8, p.hidden)) = o;
  }
}

// one left-over item cut along the sequence: every part publishes (max, sum, acc[64]); the part that arrives last
// (monotonic counter, MEGA_ATT_SPLIT arrivals per item, layer and step) merges them and writes the output row.
template <int GM, bool PROF>
__device__ void attention_part(const MegaParams& p, int layer, int bh, int pos, int extra, int q, int nsplit, int u0, int u1,
                               bool issued, float* wsm, uint8_t* ring, int nslot, uint64_t* bars, uint32_t& par) {
  const int lane = threadIdx.x & 31;
  float m, l, acc[8];
  attention_stream<GM, PROF>(p, layer, bh, pos, u0, u1, q == nsplit - 1, issued, wsm, ring, nslot, bars, par, m, l, acc);
  float* rec = p.attn_part + ((size_t)extra * nsplit + q) * 72;
  if ((lane >> 3) == 0) {
    float4* dst = reinterpret_cast<float4*>(rec + 8 + (lane & 7) * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (lane == 0) { rec[0] = m; rec[1] = l; }
  }
  __threadfence();                    // every writer releases its part of the record, then lane 0 counts the arrival
  __syncwarp();
  unsigned int old = 0;
  if (lane == 0) {
    old = atomicAdd(p.attn_cnt + extra, 1u);
    __threadfence();
  }
  old = __shfl_sync(0xffffffffu, old, 0);
  if ((old + 1u) % (unsigned int)nsplit != 0u) return;
  // last arriver: merge (records are read past L1 -- the same addresses were read one layer ago)
  const float* base = p.attn_part + (size_t)extra * nsplit * 72;
  float mm[MEGA_ATT_MAXSPLIT], M = -INFINITY;
#pragma unroll
  for (int k = 0; k < MEGA_ATT_MAXSPLIT; ++k) { mm[k] = k < nsplit ? __ldcg(base + k * 72) : -INFINITY; M = fmaxf(M, mm[k]); }
  float L = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
  for (int k = 0; k < MEGA_ATT_MAXSPLIT; ++k) {
    if (k >= nsplit) break;
    const float w = __expf(mm[k] - M);
    L = fmaf(__ldcg(base + k * 72 + 1), w, L);
    const float4 a0 = __ldcg(reinterpret_cast<const float4*>(base + k * 72 + 8 + (lane & 7) * 8));
    const float4 a1 = __ldcg(reinterpret_cast<const float4*>(base + k * 72 + 8 + (lane & 7) * 8) + 1);
    o[0] = fmaf(a0.x, w, o[0]); o[1] = fmaf(a0.y, w, o[1]); o[2] = fmaf(a0.z, w, o[2]); o[3] = fmaf(a0.w, w, o[3]);
    o[4] = fmaf(a1.x, w, o[4]); o[5] = fmaf(a1.y, w, o[5]); o[6] = fmaf(a1.z, w, o[6]); o[7] = fmaf(a1.w, w, o[7]);
  }
  attention_store(p, bh, o, 1.0f / L);
}

// (History: v3 ran 512 threads with a PAIR of warps per item, flash-decoding style halves combined through shared memory:
// 55 us per layer against 51-53 for one warp per item; removed.)

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
  constexpr int NW = MEGA_THREADS / 32;
  __shared__ float s_redf[NW];
  __shared__ int s_redi[NW];
#ifndef IVG_MEGA_NO_SLOTS
  if (p.slot_period > 0 && pos + 1 >= p.slot0 && (pos + 1 - p.slot0) % p.slot_period == 0) {
    if (tid == 0) *out = p.slot_token;       // forced separator, never sampled (action_model.py:109-110); CTA-uniform
    return;
  }
#endif
  if (!p.do_sample) {
    float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll 8
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
      for (int i = 1; i < NW; ++i) if (s_redf[i] > bv || (s_redf[i] == bv && s_redi[i] < bi)) { bv = s_redf[i]; bi = s_redi[i]; }
      *out = bi;
    }
    __syncthreads();
    return;
  }
  uint32_t* keys = smem_u;
  uint32_t* hist = smem_u + V;        // 256 bins (radix fallback) / candidate list (MEGA_SAMPLE_CAP entries)
  constexpr int CAP = 192;
  __shared__ uint32_t s_prefix, s_remaining, s_bound, s_nc;
  __shared__ __align__(16) uint32_t s_tmax[MEGA_THREADS];
  __shared__ float s_hi[MEGA_THREADS];
  __shared__ int s_win, s_lastmass;
  const uint32_t want = (uint32_t)(p.topk < V ? p.topk : V);
  uint32_t tmax = 0u;
  {   // 16-byte loads, all of a thread's requests in flight before the first use (rows are 16-byte aligned: ldl % 4 == 0)
    const int V4 = V >> 2;
    const float4* row4 = reinterpret_cast<const float4*>(row);
    uint4* keys4 = reinterpret_cast<uint4*>(keys);
#pragma unroll 4
    for (int c = tid; c < V4; c += MEGA_THREADS) {
      const float4 v = row4[c];
      const uint4 k4 = make_uint4(mega_fkey(v.x * p.inv_temp), mega_fkey(v.y * p.inv_temp), mega_fkey(v.z * p.inv_temp),
                                  mega_fkey(v.w * p.inv_temp));
      keys4[c] = k4;
      tmax = max(max(tmax, k4.x), max(k4.y, max(k4.z, k4.w)));
    }
    for (int c = (V4 << 2) + tid; c < V; c += MEGA_THREADS) { const uint32_t k = mega_fkey(row[c] * p.inv_temp); keys[c] = k; tmax = max(tmax, k); }
  }
  s_tmax[tid] = tmax;
  if (tid == 0) { s_prefix = 0; s_remaining = want; s_win = 0x7fffffff; s_lastmass = 0; s_bound = 0u; s_nc = 0u; }
  __syncthreads();
  // ---- k-th largest key.  Fast path: a lower bound from the per-thread maxima (the want-th largest of them has at
  // least `want` keys at or above it), the ~130 keys above the bound are compacted and ranked exhaustively.  The radix
  // select below remains for top_k > 256 and for degenerate rows (many equal logits). ----
  bool fast = want <= (uint32_t)MEGA_THREADS;
  if (fast) {
    uint32_t ge = 0;
    const uint4* tm4 = reinterpret_cast<const uint4*>(s_tmax);
#pragma unroll 8
    for (int j = 0; j < MEGA_THREADS / 4; ++j) {
      const uint4 t = tm4[j];
      ge += (t.x >= tmax) + (t.y >= tmax) + (t.z >= tmax) + (t.w >= tmax);
    }
    uint32_t cand_b = ge >= want ? tmax : 0u;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) cand_b = max(cand_b, __shfl_xor_sync(0xffffffffu, cand_b, off));
    if (lane == 0 && cand_b) atomicMax(&s_bound, cand_b);
    __syncthreads();
    const uint32_t bound = s_bound;
    for (int c = tid; c < V; c += MEGA_THREADS) {
      const uint32_t kk = keys[c];
      if (kk >= bound) { const uint32_t i = atomicAdd(&s_nc, 1u); if (i < (uint32_t)CAP) hist[i] = kk; }
    }
    __syncthreads();
    const uint32_t nc = s_nc;
    fast = nc <= (uint32_t)CAP;
    if (fast) {
      uint32_t best = 0u;
      for (uint32_t c = tid; c < nc; c += MEGA_THREADS) {
        const uint32_t mine = hist[c];
        uint32_t cnt = 0;
        for (uint32_t j = 0; j < nc; ++j) cnt += hist[j] >= mine;
        if (cnt >= want) best = max(best, mine);
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
      if (lane == 0 && best) atomicMax(&s_prefix, best);
    }
    __syncthreads();
  }
  if (!fast) {
    if (tid == 0) s_prefix = 0;
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 256; i += MEGA_THREADS) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
      // logits share a few exponents, so the leading digits fall into a handful of bins: aggregate equal bins inside
      // the warp (match.any) and let one lane add the count
      for (int c0 = 0; c0 < V; c0 += MEGA_THREADS) {
        const int c = c0 + tid;
        const uint32_t kk = c < V ? keys[c] : 0u;
        const bool in = c < V && (kk & mask) == prefix;
        const uint32_t bin = in ? ((kk >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
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
  }
  const uint32_t kth = s_prefix;
  uint32_t mk = tmax;                 // this thread's keys were all loaded by itself above
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mk = max(mk, __shfl_xor_sync(0xffffffffu, mk, off));
  if (lane == 0) s_redi[warp] = (int)mk;
  __syncthreads();
  for (int i = 0; i < NW; ++i) mk = max(mk, (uint32_t)s_redi[i]);
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
  for (int i = 0; i < NW; ++i) { const float v = s_redf[i]; if (i < warp) wbase += v; total += v; }
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

template <int GM, int AM, bool PROF>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(const MegaParams p) {
  constexpr int MAXP = GM == 0 ? 8 : MEGA_MAX_SPLITS;
  extern __shared__ uint8_t mega_raw[];
  MegaCtx c;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mega_raw) + 1023) & ~(uintptr_t)1023);
  c.sm.a = base;
  if constexpr (GM == 0) {   // activation slab sized for this model, the rest of the 192 KB operand area holds 2-4 weight slab buffers
    const int a_rows = p.B <= 64 ? 64 : 128;
    const int kd = p.inter / p.d_splits;
    const int kmax = p.hidden > kd ? p.hidden : kd;
    // narrow phases: qkv / o-proj items are 16 rows x <= hidden, down-proj items bn_down rows x inter / d_splits
    const uint32_t s_qkv = (uint32_t)(MEGA_BN * p.hidden * 2), s_down = (uint32_t)(p.bn_down * kd * 2);
    c.sm.slab_bytes = s_qkv > s_down ? s_qkv : s_down;
    c.sm.a_bytes = (uint32_t)(a_rows * kmax * 2);
    // floor: the attention ring and the sampler scratch live in this region too -- as much of 128 KB as leaves two slabs
    uint32_t floor_b = (uint32_t)(MEGA_A_BYTES + 2 * MEGA_B_BYTES) - 2u * c.sm.slab_bytes;
    floor_b = floor_b > (uint32_t)MEGA_A_BYTES ? (uint32_t)MEGA_A_BYTES : floor_b;
    if (c.sm.a_bytes < floor_b) c.sm.a_bytes = floor_b;
    const uint32_t nb = (uint32_t)(MEGA_A_BYTES + 2 * MEGA_B_BYTES - c.sm.a_bytes) / c.sm.slab_bytes;
    c.sm.nbuf = nb > 4u ? 4u : nb;
    c.sm.b0 = base + c.sm.a_bytes;
  } else {                  // weight-stationary mode: activations of the widest K range (gate/up, lm_head: hidden) + 32 KB weight slabs
    const int kd = p.inter / p.d_splits;
    const int kmax = p.hidden > kd ? p.hidden : kd;
    c.sm.a_bytes = (uint32_t)(p.a_rows * kmax * 2);
    // floor: the attention ring lives here too and streams faster with more slots (96 KB gave 4 slots per warp and a
    // 15 % slower K/V loop than the 128 KB = 5 slots of mode 0, profiles/r01 v12)
    if (c.sm.a_bytes < 128u * 1024u) c.sm.a_bytes = 128u * 1024u;
    c.sm.slab_bytes = (uint32_t)(MEGA_WM * MEGA_W_CHUNK * 2);
    const uint32_t nb = (uint32_t)(MEGA_A_BYTES + 2 * MEGA_B_BYTES - c.sm.a_bytes) / c.sm.slab_bytes;
    c.sm.nbuf = nb > 4u ? 4u : nb;
    c.sm.b0 = base + c.sm.a_bytes;
  }
  c.sm.bfull = reinterpret_cast<uint64_t*>(base + MEGA_A_BYTES + 2 * MEGA_B_BYTES);
  c.sm.mma_done = c.sm.bfull + 4;
  c.sm.tmem_holder = reinterpret_cast<uint32_t*>(c.sm.mma_done + 1);
  c.sm.abar = c.sm.bfull + 8;
  c.sm.bempty = c.sm.bfull + 9;
  c.sm.ring_bar = c.sm.bfull + 16;                                   // 64 barriers, 128 B past the GEMM ones
  c.sm.sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(c.sm.bfull) + MEGA_BAR_BYTES);
  uint32_t ring_par = 0;                                             // expected parity per ring slot of this warp
  c.epoch = 0; c.issued = 0; c.consumed = 0; c.phase_issued = 0; c.phase_consumed = 0; c.wait_par = 0; c.mphase = 0; c.aphase = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(c.sm.bfull + i, 1);
    mbar_init(c.sm.mma_done, GM == 0 ? MEGA_NISSUE : 1);   // gemm_mode 0: one commit per issuing warp
    mbar_init(c.sm.abar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(c.sm.bempty + i, 1);
    for (int i = 0; i < 64; ++i) mbar_init(c.sm.ring_bar + i, 1);
    fence_barrier_init();
  }
  constexpr uint32_t TMEM_COLS = GM == 0 ? (MEGA_NACC <= 1 ? 64 : (MEGA_NACC == 2 ? 128 : 256)) : 128;   // mode 0: MEGA_NACC accumulators of <= 64 columns; mode 1: a_rows <= 128
  if (warp == 1) { tmem_alloc(c.sm.tmem_holder, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *c.sm.tmem_holder;

  const int H = p.hidden;
  const int pos0 = *p.dpos;
  uint32_t* smem_u = reinterpret_cast<uint32_t*>(c.sm.a);
  bool ok = true;
  // phase timing (CTA 0, thread 0): slots 0 norm, 1 qkv, 2 attention, 3 o-proj, 4 gate/up, 5 down, 6 lm_head,
  // 7 sample, 8 barriers
  long long tprof[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool profiling = PROF && p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long tmark = clock64();
#define MEGA_MARK(slot) do { if (profiling) { const long long _t = clock64(); tprof[slot] += _t - tmark; tmark = _t; } } while (0)
#define MEGA_BARRIER(F) do { ok = grid_barrier<F>(p, c); MEGA_MARK(8); } while (0)

  for (int step = 0; step < p.steps && ok; ++step) {
    const int pos = pos0 + step;
    constexpr bool ws = GM != 0;            // weight-stationary GEMM phases: qkv goes through split-K fp32 partials
    // gemm_mode 0 slab geometry: narrow phases (qkv, o, down: 16-row items, one round over the SMs) keep their buffers above
    // the largest activation slab; the wide phases (gate/up, lm_head: K = hidden) put theirs right above the K = hidden slab
    const uint32_t narrow_off = c.sm.a_bytes, narrow_bytes = c.sm.slab_bytes, narrow_nbuf = c.sm.nbuf;
    const int a_rows0 = p.B <= 64 ? 64 : 128;
    const uint32_t wide_off = (uint32_t)(a_rows0 * H * 2), wide_bytes = (uint32_t)(p.bn_wide * H * 2);
    uint32_t wide_nbuf = (uint32_t)(MEGA_A_BYTES + 2 * MEGA_B_BYTES - wide_off) / wide_bytes;
    wide_nbuf = wide_nbuf > 4u ? 4u : wide_nbuf;
    GemmPhase qkv_g{p.lw[0].wqkv, 3 * H, H, ws ? p.qkv_splits : 1, p.xn, H, ws ? EPI_PARTIAL_F32 : EPI_STORE_BF16,
                    ws ? (void*)p.qkvp : (void*)p.qkv, 3 * H, MEGA_BN, narrow_off, narrow_bytes, narrow_nbuf};
    constexpr bool FN = GM == 0 && MEGA_FUSE_NORM;                          // fused norms (decode_mega.cuh)
    qkv_g.norm_a = FN ? 1 : 0;
    prefetch_phase<GM>(p, c, qkv_g);
    norm_phase<MAXP>(p, p.lw[0].n1, 0, p.tokens, pos, FN);                   // x = E[token], xn = rmsnorm(x) (FN: bf16(x))
    MEGA_MARK(0);
    MEGA_BARRIER(true); if (!ok) break;
    for (int l = 0; l < p.layers && ok; ++l) {
      const MegaLayer& L = p.lw[l];
      qkv_g.w = L.wqkv;
      gemm_phase<GM, PROF>(p, c, qkv_g);
      MEGA_MARK(1);
      GemmPhase o_g{L.wo, H, H, p.o_splits, p.ao, H, EPI_PARTIAL_F32, p.part, H, MEGA_BN, narrow_off, narrow_bytes, narrow_nbuf};
      if (FN) o_g.resid_cnt = p.tile_cnt;
      prefetch_phase<GM>(p, c, o_g);
      const bool ring_prefetch = AM == 0 && p.attn_mode == 0 && att_even_deal(p);
      if (ring_prefetch) {
        // the activation slab is dead (this CTA's MMAs have retired): start filling the attention ring with old K/V rows.
        // The region was last written through the generic proxy (cp.async), the copies below are async-proxy writes.
        fence_proxy_async();
        const AttWork w = att_assign(p, c.sm.a_bytes, warp, pos);
        attention_prefetch(p, l, pos, w, c.sm.a + (size_t)warp * w.nslot * MEGA_RING_SLOT, c.sm.ring_bar + warp * 8);
      }
      MEGA_BARRIER(false); if (!ok) break;
      static_assert(MEGA_THREADS == 256, "the attention phase deals one warp per item over 8 warps");
      {
        if constexpr (AM != 1) {
          // the ring lives in the activation region, last written through the generic proxy (cp.async / scratch)
          fence_proxy_async();
          const int total = p.B * p.heads, G = (int)gridDim.x;
          if (att_even_deal(p)) {
            const AttWork w = att_assign(p, c.sm.a_bytes, warp, pos);     // same assignment the prefetch used
            float* wsm = c.sm.sc + (size_t)warp * (p.Lmax + 8 + 64);
            uint8_t* ring = c.sm.a + (size_t)warp * w.nslot * MEGA_RING_SLOT;
            if (w.kind == 1) {
              float m, lsum, acc[8];
              attention_stream<GM, PROF>(p, l, w.bh, pos, w.u0, w.u1, true, ring_prefetch, wsm, ring, w.nslot, c.sm.ring_bar + warp * 8,
                               ring_par, m, lsum, acc);
              attention_store(p, w.bh, acc, 1.0f / lsum);
            } else if (w.kind == 2) {
              attention_part<GM, PROF>(p, l, w.bh, pos, w.extra, w.q, w.nsplit, w.u0, w.u1, ring_prefetch, wsm, ring, w.nslot,
                             c.sm.ring_bar + warp * 8, ring_par);
            }
          } else {
            for (int base = blockIdx.x; base < total; base += G * 8) {
              int active = (total - base + G - 1) / G;            // warps of this CTA with an item this round
              active = active > 8 ? 8 : active;
              int nslot = (int)(c.sm.a_bytes / MEGA_RING_SLOT) / active;
              nslot = nslot > 8 ? 8 : nslot;
              if (warp < active) {
                const int bh = base + G * warp;
                float m, lsum, acc[8];
                attention_stream<GM, PROF>(p, l, bh, pos, 0, (pos + 31) >> 5, true, false, c.sm.sc + (size_t)warp * (p.Lmax + 8 + 64),
                                 c.sm.a + (size_t)warp * nslot * MEGA_RING_SLOT, nslot, c.sm.ring_bar + warp * 8, ring_par, m,
                                 lsum, acc);
                attention_store(p, bh, acc, 1.0f / lsum);
              }
              __syncthreads();                 // the next round partitions the ring differently
            }
          }
        } else {
          for (int bh = blockIdx.x + (int)gridDim.x * warp; bh < p.B * p.heads; bh += (int)gridDim.x * (MEGA_THREADS / 32))
            attention_warp<GM>(p, l, bh, pos, c.sm.sc + (size_t)warp * (p.Lmax + 8 + 64));
        }
      }
      MEGA_MARK(2);
      MEGA_BARRIER(true); if (!ok) break;
      gemm_phase<GM, PROF>(p, c, o_g);
      MEGA_MARK(3);
      GemmPhase gu_g{L.wgu, 2 * p.inter, H, 1, p.xn, H, EPI_SWIGLU, p.act, p.inter, p.bn_wide, wide_off, wide_bytes, wide_nbuf};
      gu_g.norm_a = FN ? 1 : 0;
      prefetch_phase<GM>(p, c, gu_g);
      if constexpr (!FN) {
        MEGA_BARRIER(false); if (!ok) break;
        norm_phase<MAXP>(p, L.n2, p.o_splits, nullptr, 0);                  // x += o partials; xn = rmsnorm(x) * n2
        MEGA_MARK(0);
      }
      MEGA_BARRIER(true); if (!ok) break;
      gemm_phase<GM, PROF>(p, c, gu_g);
      MEGA_MARK(4);
      GemmPhase d_g{L.wd, H, p.inter, p.d_splits, p.act, p.inter, EPI_PARTIAL_F32, p.part, H, p.bn_down, narrow_off, narrow_bytes,
                    narrow_nbuf};
      if (FN) d_g.resid_cnt = p.tile_cnt + 64;
      prefetch_phase<GM>(p, c, d_g);
      MEGA_BARRIER(true); if (!ok) break;
      gemm_phase<GM, PROF>(p, c, d_g);
      MEGA_MARK(5);
      const bool last = (l == p.layers - 1);
      GemmPhase nx_g{last ? p.lm_head : p.lw[l + 1].wqkv, last ? p.vocab : 3 * H, H, (ws && !last) ? p.qkv_splits : 1, p.xn, H,
                     last ? EPI_LOGITS : (ws ? EPI_PARTIAL_F32 : EPI_STORE_BF16),
                     last ? (void*)p.logits : (ws ? (void*)p.qkvp : (void*)p.qkv), last ? p.ldl : (long long)(3 * H),
                     last ? p.bn_wide : MEGA_BN, last ? wide_off : narrow_off, last ? wide_bytes : narrow_bytes,
                     last ? wide_nbuf : narrow_nbuf};
      nx_g.norm_a = FN ? 1 : 0;
      prefetch_phase<GM>(p, c, nx_g);
      if constexpr (!FN) {
        MEGA_BARRIER(false); if (!ok) break;
        norm_phase<MAXP>(p, last ? p.norm_f : p.lw[l + 1].n1, p.d_splits, nullptr, 0);   // x += down partials; next norm
        MEGA_MARK(0);
      }
      MEGA_BARRIER(true); if (!ok) break;
      if (last) {
        gemm_phase<GM, PROF>(p, c, nx_g);                                       // lm_head
        MEGA_MARK(6);
        MEGA_BARRIER(false); if (!ok) break;
        for (int b = blockIdx.x; b < p.B; b += gridDim.x) sample_row(p, b, pos, smem_u);
        MEGA_MARK(7);
        MEGA_BARRIER(false); if (!ok) break;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && ok) *p.dpos = pos0 + p.steps;
  if (profiling) for (int i = 0; i < 9; ++i) p.prof[i] += tprof[i];   // slots 9..13: attention_stream's own marks
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(c.tmem_base, TMEM_COLS); }
}

// ---- one-off weight packing: [rows, cols] bf16 row-major -> per 16-row work item the 128B-swizzled K-major image ----
// out[tile][k-block][row < bn][chunk' = chunk ^ (row & 7)][8]; rows padded with zeros to a multiple of bn (the phase's tile width).
__global__ void mega_pack_weight_kernel(const uint4* __restrict__ w, uint4* __restrict__ out, int rows, int cols, int bn,
                                        long long chunks) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i & 7);
    const int r = (int)((i >> 3) % bn);
    const long long jt = (i >> 3) / bn;
    const int nkb = cols >> 6;
    const int j = (int)(jt % nkb);
    const long long tile = jt / nkb;
    const long long n = tile * bn + r;
    const int ch = cp ^ (r & 7);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < rows) v = w[(n * cols + j * 64 + ch * 8) >> 3];
    out[i] = v;
  }
}

int mega_pack_weight_launch(const void* w, void* out, int rows, int cols, int bn, cudaStream_t st) {
  IVG_CHECK(cols % 64 == 0 && rows >= 1, "mega_pack_weight: cols %d must be a multiple of 64", cols);
  IVG_CHECK(bn >= 16 && bn <= 256 && bn % 16 == 0, "mega_pack_weight: tile width %d must be a multiple of 16 in [16, 256]", bn);
  const long long tiles = (rows + bn - 1) / bn;
  const long long chunks = tiles * (cols / 64) * bn * 8;
  const int blocks = (int)((chunks + 255) / 256 < 148 * 16 ? (chunks + 255) / 256 : 148 * 16);
  mega_pack_weight_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(w), reinterpret_cast<uint4*>(out), rows,
                                                  cols, bn, chunks);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---- weight packing for gemm_mode 1: [rows, cols] bf16 row-major -> per 64-row tile the 128B-swizzled K-major image ----
// out[tile][k-block][R (64)][chunk ^ (R & 7)][8]; rows padded with zeros to a multiple of 64.  swiglu_pairs: the source is the
// row-interleaved (gate_0, up_0, gate_1, up_1, ...) matrix; tile t covers outputs [32 t, 32 t + 32) and each 16-row group g
// holds gate rows of outputs 32 t + 8 g + (0..7) followed by their up rows, so the pair sits 8 TMEM lanes apart in one warp.
__global__ void mega_pack_weight64_kernel(const uint4* __restrict__ w, uint4* __restrict__ out, int rows, int cols,
                                          int swiglu_pairs, long long chunks) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i & 7);
    const int R = (int)((i >> 3) & 63);
    const long long jt = i >> 9;
    const int nkb = cols >> 6;
    const int j = (int)(jt % nkb);
    const long long tile = jt / nkb;
    long long n = tile * 64 + R;
    if (swiglu_pairs) {
      const int g = R >> 4, r = R & 15;
      n = 2 * (tile * 32 + 8 * g + (r & 7)) + (r >> 3);
    }
    const int ch = cp ^ (R & 7);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < rows) v = w[(n * cols + j * 64 + ch * 8) >> 3];
    out[i] = v;
  }
}

int mega_pack_weight64_launch(const void* w, void* out, int rows, int cols, int swiglu_pairs, cudaStream_t st) {
  IVG_CHECK(cols % 64 == 0 && rows >= 1, "mega_pack_weight64: cols %d must be a multiple of 64", cols);
  IVG_CHECK(!swiglu_pairs || rows % 2 == 0, "mega_pack_weight64: gate/up interleaved matrix needs an even row count");
  const long long tiles = (rows + MEGA_WM - 1) / MEGA_WM;
  const long long chunks = tiles * (cols / 64) * 64 * 8;
  const int blocks = (int)((chunks + 255) / 256 < 148 * 16 ? (chunks + 255) / 256 : 148 * 16);
  mega_pack_weight64_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(w), reinterpret_cast<uint4*>(out), rows,
                                                    cols, swiglu_pairs, chunks);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int mega_fused_norm() { return MEGA_FUSE_NORM ? 1 : 0; }

int decode_mega_launch(const MegaParams& p, int num_sms, cudaStream_t st) {
  IVG_CHECK(!(MEGA_FUSE_NORM && p.gemm_mode == 0) || (p.tile_cnt != nullptr && p.hidden / MEGA_BN <= 64 && p.o_splits <= 8 &&
                                                      p.d_splits <= 8),
            "decode_mega: fused norms need tile_cnt [128 uints, zeroed], hidden <= 1024 and at most 8 splits");
  IVG_CHECK(p.B >= 1 && p.B <= 128, "decode_mega: batch %d not in [1,128]", p.B);
  IVG_CHECK(p.hidden % 64 == 0 && p.hidden <= MEGA_MAXK, "decode_mega: hidden %d unsupported", p.hidden);
  IVG_CHECK(p.gemm_mode == 0 || p.gemm_mode == 1, "decode_mega: gemm_mode %d", p.gemm_mode);
  const int max_splits = p.gemm_mode == 0 ? 8 : MEGA_MAX_SPLITS;
  IVG_CHECK(p.o_splits >= 1 && p.o_splits <= max_splits && p.hidden % (64 * p.o_splits) == 0,
            "decode_mega: bad o_splits %d for hidden %d", p.o_splits, p.hidden);
  IVG_CHECK(p.d_splits >= 1 && p.d_splits <= max_splits && p.inter % (64 * p.d_splits) == 0 &&
                (p.gemm_mode != 0 || p.inter / p.d_splits <= MEGA_MAXK),
            "decode_mega: bad d_splits %d for intermediate size %d", p.d_splits, p.inter);
  IVG_CHECK(p.hidden == p.heads * 64, "decode_mega: head_dim must be 64");
  if (p.gemm_mode == 0) {
    const int a_rows = p.B <= 64 ? 64 : 128;
    IVG_CHECK(p.a_rows == a_rows, "decode_mega: a_rows must be %d in gemm_mode 0", a_rows);
    IVG_CHECK((long long)a_rows * p.hidden * 2 <= MEGA_A_BYTES && (long long)a_rows * (p.inter / p.d_splits) * 2 <= MEGA_A_BYTES,
              "decode_mega: batch %d with K %d does not fit the shared-memory activation slab", p.B, p.hidden);
    IVG_CHECK(p.bn_down >= 16 && p.bn_down <= 64 && p.bn_down % 16 == 0 &&
                  (long long)a_rows * (p.inter / p.d_splits) * 2 + 2ll * p.bn_down * (p.inter / p.d_splits) * 2 <=
                      MEGA_A_BYTES + 2 * MEGA_B_BYTES,
              "decode_mega: down-proj tile width %d with %d splits does not leave two weight slabs", p.bn_down, p.d_splits);
    IVG_CHECK(p.bn_wide >= 16 && p.bn_wide <= 64 && p.bn_wide % 16 == 0 &&
                  (long long)(a_rows + p.bn_wide) * p.hidden * 2 <= MEGA_A_BYTES + 2 * MEGA_B_BYTES,
              "decode_mega: wide tile width %d (gate/up, lm_head) must be a multiple of 16 in [16, 64] whose slab fits next to the "
              "activation slab", p.bn_wide);
  } else {
    IVG_CHECK(p.a_bulk == 1, "decode_mega: gemm_mode 1 needs the swizzled activation images (a_bulk)");
    IVG_CHECK(p.a_rows >= p.B && p.a_rows % 8 == 0 && p.a_rows <= 128, "decode_mega: a_rows %d must be a multiple of 8 in [B, 128]",
              p.a_rows);
    IVG_CHECK(p.qkv_splits >= 1 && p.qkv_splits <= MEGA_MAX_QKV_SPLITS && p.hidden % (64 * p.qkv_splits) == 0 && p.qkvp != nullptr,
              "decode_mega: bad qkv_splits %d for hidden %d", p.qkv_splits, p.hidden);
    const int kd = p.inter / p.d_splits;
    const long long a_bytes = (long long)p.a_rows * (p.hidden > kd ? p.hidden : kd) * 2;
    const long long a_floor = a_bytes < 128 * 1024 ? 128 * 1024 : a_bytes;
    IVG_CHECK(a_floor + 2 * MEGA_WM * MEGA_W_CHUNK * 2 <= MEGA_A_BYTES + 2 * MEGA_B_BYTES,
              "decode_mega: batch %d with K %d leaves no room for two weight slabs", p.B, p.hidden);
  }
  IVG_CHECK((size_t)(p.Lmax + 8 + 64) * 4 * (MEGA_THREADS / 32) <= MEGA_SC_BYTES && (size_t)(p.vocab + 256) * 4 <= MEGA_A_BYTES,
            "decode_mega: Lmax/vocab too large for the scratch region");
  IVG_CHECK(p.Lmax % 8 == 0, "decode_mega: Lmax must be a multiple of 8");
  // one instantiation per (GEMM mode, attention mode, profiling): the timed kernels carry no timing code at all (cycle
  // counters in the 250-register kernel cost 4.5 ms per rollout even when disabled at run time, same-box A/B)
  const void* fn = nullptr;
  const int am = p.attn_mode == 1 ? 1 : 0;
  const int pf = p.prof != nullptr ? 1 : 0;
  const void* table[8] = {
      (const void*)decode_mega_kernel<0, 0, false>, (const void*)decode_mega_kernel<0, 0, true>,
      (const void*)decode_mega_kernel<0, 1, false>, (const void*)decode_mega_kernel<0, 1, true>,
      (const void*)decode_mega_kernel<1, 0, false>, (const void*)decode_mega_kernel<1, 0, true>,
      (const void*)decode_mega_kernel<1, 1, false>, (const void*)decode_mega_kernel<1, 1, true>};
  const int vi = (p.gemm_mode != 0 ? 4 : 0) + am * 2 + pf;
  fn = table[vi];
  static PerDeviceOnce attr_once[8];
  if (attr_once[vi].pending()) {
    IVG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_SMEM));
    attr_once[vi].mark();
  }
  int max_blocks = 0;
  IVG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, fn, MEGA_THREADS, MEGA_SMEM));
  IVG_CHECK(max_blocks >= 1, "decode_mega: kernel does not fit on an SM");
  void* args[] = {const_cast<MegaParams*>(&p)};
  IVG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(num_sms), dim3(MEGA_THREADS), args, (size_t)MEGA_SMEM, st));
  count_launch();
  return 0;
}

}  // namespace ivg
