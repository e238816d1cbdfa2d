// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )          fp32 accumulation in TMEM
//
// One kernel covers every dense contraction on the iVideoGPT hot path:
//   * mode 0 (plain / batched): Llama q/k/v/o/gate/up/down/lm_head projections
//     (transformers LlamaAttention / LlamaMLP, called from reference inference/predict.py:64 via generate),
//     1x1 convs and Linear layers of the tokenizer (compressive_vq_model.py:188,196,241,245),
//     Q.K^T and P.V of the cross/self attention blocks (conditional_vae.py:49), with per-head column offsets
//     so that no head split/merge copies are ever made.
//   * mode 1 (conv taps): 3x3 convolutions of the ctx_vqgan encoder/decoder (vae.py:86-137,236-294) as an
//     implicit GEMM over NHWC activations: the A operand of k-block (tap, channel-chunk) is a 4-D TMA box
//     of th x tw output pixels shifted by the tap; TMA out-of-bounds zero fill *is* the conv padding.
//     Stride-2 convs (diffusers Downsample2D, pad (0,1,0,1)) read one of four parity views of the input.
//     A trailing "extra" segment of k-blocks reads a second tensor at tap (0,0): the fused 1x1 shortcut of
//     ResnetBlock2D.
//
// Structure (per CTA, 192 threads, 1 CTA/SM, grid = min(#tiles, #SMs), static tile striding):
//   warp 0 (one elected lane) : TMA producer       -> smem ring of S stages {A 128x128B, B BNx128B}, SWIZZLE_128B
//   warp 1 (one elected lane) : tcgen05.mma issuer -> 2 TMEM accumulator stages of BN fp32 columns each
//   warp 6 (one elected lane) : SECOND tcgen05.mma issuer.  One thread issuing back to back sustains one 128 x 256 x 16 MMA per
//                               ~160 cycles although the tensor core needs 128 (tools/probes/mma_probe.cu,
//                               profiles/r02/mma_issue_and_ffma_probe.json): a single-issuer kernel is capped at 80 % of peak.
//                               Two warps issuing ALTERNATE k-blocks into the SAME accumulator reach 127 cycles per MMA with exact
//                               results (the probe checks every element).  The order in which the tensor core retires the two
//                               streams is not fixed, so the fp32 summation order -- the last bits of the result -- may differ
//                               from run to run.  OPT-IN (GemmParams::issuers = 2, IVGPT_MMA_ISSUERS=2): in this kernel the
//                               limit is operand supply from L2, the second issuer changed nothing (46.6 vs 46.8 ms of conv time
//                               per cfg64 step), so the default stays the single bit-reproducible issuer.
//   warps 2-5                 : epilogue, tcgen05.ld 32 lanes x 16 columns, bias/residual/activation, global stores
// Operands are bf16 (kind::f16) or fp32 read as tf32 (kind::tf32); a k-block is always 128 bytes of K.
#include <type_traits>
#include <utility>
#include <vector>

#include "common.cuh"
#include "gemm_params.cuh"

namespace ivg {

// MH = M sub-tiles of 128 rows per CTA tile.  MH = 2 (with BN = 256): a 256 x 256 output tile per CTA, both halves multiply the
// SAME weight tile, so a k-block moves A 2 x 16 KB + B 32 KB for 2 x (128 x 256 x 64) MACs: 64 B/clk/SM instead of 96 -- the
// kernel is bound by operand supply from L2, not by the tensor pipe (a second MMA issuer changed nothing, profiles/r02).  The
// two halves own the two TMEM accumulators, so the accumulator is not double-buffered across tiles any more: the epilogue of
// the second half is exposed once per tile (K >= 1152 for the convs: a few percent).
template <int BN, int MH = 1>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_ROWB;
  static constexpr int B_BYTES = BN * GEMM_ROWB;
  static constexpr int STAGE_BYTES = MH * A_BYTES + B_BYTES;
  static constexpr int STAGES = MH == 2 ? 3 : ((BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8));
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int GN_OFFSET = BAR_OFFSET + 256;     // 4 warps x 32 groups x (sum, sumsq) fp32 = 1 KB
  // fast-epilogue GroupNorm statistics: per-row partials [128 rows][17 (16 slots, padded)] float2, per-16-row-block sums
  // [8][16] float2, per-tile group sums [32] float2
  static constexpr int GNR_OFFSET = GN_OFFSET + 1024;
  static constexpr int GNR_BYTES = 128 * 17 * 8 + 8 * 16 * 8 + 32 * 8;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024 + GNR_BYTES + 1024;  // + barrier block + GN scratch + alignment slack
};

// ---- operand transform: GroupNorm affine + SiLU on 8 bf16 / 4 fp32 channels of one pixel (one 16-byte chunk) ----
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ float tanh_f32(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// silu(y) = y * sigmoid(y) = h + h * tanh(h) with h = y / 2: ONE special-function op per element (per PAIR in bf16x2), where
// y / (1 + exp(-y)) costs two -- the transform of a k-block must stay under the ~512 tensor cycles the same k-block takes
template <typename T> struct XfChunk;
template <> struct XfChunk<__nv_bfloat16> {
  static constexpr int NCH = 8;
  __device__ static __forceinline__ uint4 apply(uint4 v, const float* sc, const float* sh, bool silu) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      const float2 f = __bfloat1622float2(x2);
      const float y0 = fmaf(f.x, sc[2 * i], sh[2 * i]), y1 = fmaf(f.y, sc[2 * i + 1], sh[2 * i + 1]);
      if (silu) {                                 // the caller passes HALVED coefficients: y0 / y1 are already h = y / 2
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const uint32_t tb = tanh_bf16x2(hb);
        const __nv_bfloat162 o2 = __hfma2(h2, *reinterpret_cast<const __nv_bfloat162*>(&tb), h2);
        w[i] = *reinterpret_cast<const uint32_t*>(&o2);
      } else {
        w[i] = pack_bf16x2(y0, y1);
      }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct XfChunk<float> {
  static constexpr int NCH = 4;
  __device__ static __forceinline__ uint4 apply(uint4 v, const float* sc, const float* sh, bool silu) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float y = fmaf(__uint_as_float(w[i]), sc[i], sh[i]);
      if (silu) y = fmaf(y, tanh_f32(y), y);      // halved coefficients: y is h = (x * scale + shift) / 2
      w[i] = __float_as_uint(y);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// ---- fast epilogue group: GC consecutive accumulator columns of one row ----
// The round-1 epilogue handled 16 columns at a time, strictly one after the other: tcgen05.ld -> wait -> residual load from
// global memory (~1 us of latency, dependent) -> store.  For the low-channel / high-resolution convs (K = 1152, N = 128: 4.6k
// tensor cycles per tile) that chain, not the tensor pipe, set the pace: 353 TFLOP/s on the 128 -> 128 convs at 64x64 and 256x256
// (profiles/r02/conv_shape_times_v1_before_epilogue.json).  Here a group's residual is requested BEFORE its accumulator is
// read (the first group's before the accumulator is even complete), the next group's while this one is being written, and
// the TMEM loads of a group are issued together with one wait.
template <int GC, bool RES_BF16>
struct EpiRes {                       // raw residual of one group: GC columns
  static constexpr int NV = RES_BF16 ? GC / 8 : GC / 4;       // 16-byte vectors
  uint4 v[NV];
  __device__ __forceinline__ void load(const void* rp) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(rp) + i);
  }
  __device__ __forceinline__ float get(int j) const {          // j: compile-time after unrolling
    if (RES_BF16) {
      const uint32_t w = reinterpret_cast<const uint32_t*>(v)[j >> 1];
      return (j & 1) ? __uint_as_float(w & 0xFFFF0000u) : __uint_as_float(w << 16);
    }
    return __uint_as_float(reinterpret_cast<const uint32_t*>(v)[j]);
  }
};

template <typename T, int BN, bool XF, int MH>
__global__ void __launch_bounds__(XF ? GEMM_THREADS_XF : GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
  static_assert(MH == 1 || (MH == 2 && !XF), "two M sub-tiles per CTA tile: plain operands only");
  using SM = GemmSmem<BN, MH>;
  constexpr bool TF32 = (sizeof(T) == 4);
  constexpr int BK = GEMM_ROWB / (int)sizeof(T);  // elements of K per k-block
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  constexpr uint32_t IDESC = umma_idesc(TF32 ? 2 : 1, GEMM_BM, BN);

  extern __shared__ uint8_t gemm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + SM::STAGES;
  uint64_t* tfull_bar = empty_bar + SM::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* xf_bar = tempty_bar + 2;                      // [STAGES] A tile transformed (XF): 128 arrivals
  uint64_t* start_bar = xf_bar + SM::STAGES;              // [2] the accumulator-zeroing first k-block of a tile has retired
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(start_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < SM::STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, (uint32_t)p.issuers); mbar_init(tempty_bar + s, 4); mbar_init(start_bar + s, 1); }
    if (XF) for (int s = 0; s < SM::STAGES; ++s) mbar_init(xf_bar + s, 32 * GEMM_XF_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_holder, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();               // everything above overlaps the predecessor kernel's tail under PDL
  pdl_launch_dependents();

  const int tiles_m2 = (p.tiles_m + MH - 1) / MH;          // CTA tiles along M (MH sub-tiles of 128 rows each)
  const int tiles_mn = tiles_m2 * p.tiles_n;
  const int total_tiles = tiles_mn * p.batch;
  const int num_kb = p.num_kb;

  if (warp == 0) {
    if (lane == 0) {
      // =============================== TMA producer ===============================
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int bz = tile / tiles_mn;
        const int rem = tile - bz * tiles_mn;
        const int tm2 = rem / p.tiles_n, tn = rem - tm2 * p.tiles_n;
        const int n0 = tn * BN;
        if (p.causal_skip && n0 > tm2 * GEMM_BM + GEMM_BM - 1) continue;      // (causal_skip launches use MH == 1)
        const int outer = bz / p.heads, h = bz - outer * p.heads;
        const int a_b = p.a_bsel == 0 ? 0 : (p.a_bsel == 1 ? outer / p.a_bdiv : bz);
        const int b_b = p.b_bsel == 0 ? 0 : (p.b_bsel == 1 ? outer / p.b_bdiv : bz);
        const int a_k0 = p.a_kbase + h * p.a_khead, b_k0 = p.b_kbase + h * p.b_khead;
        const int b_n0 = n0 + h * p.b_nhead;
        int m0[MH], img[MH], y0[MH], x0[MH];
#pragma unroll
        for (int hm = 0; hm < MH; ++hm) {
          const int tm = tm2 * MH + hm;                      // a sub-tile past the last one reads out of bounds: zero-filled
          m0[hm] = tm * GEMM_BM; img[hm] = 0; y0[hm] = 0; x0[hm] = 0;
          if (p.mode == 1) {
            const int tiles_x = p.W / p.tw, tiles_img = tiles_x * (p.H / p.th);
            img[hm] = tm / tiles_img;
            const int r2 = tm - img[hm] * tiles_img;
            y0[hm] = (r2 / tiles_x) * p.th;
            x0[hm] = (r2 % tiles_x) * p.tw;
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * SM::STAGE_BYTES;
          uint8_t* sb = sa + MH * SM::A_BYTES;
          mbar_expect_tx(full_bar + stage, SM::STAGE_BYTES);
#pragma unroll
          for (int hm = 0; hm < MH; ++hm) {
            uint8_t* sah = sa + hm * SM::A_BYTES;
            if (p.mode == 0) {
              tma_load_3d(sah, &maps.a[0], full_bar + stage, a_k0 + kb * BK, m0[hm], a_b);
            } else {
              int tap = kb / p.cpb;
              int chunk = kb - tap * p.cpb;
              if (tap < p.ntaps) {
                tma_load_4d(sah, &maps.a[p.tap_map[tap]], full_bar + stage, chunk * BK, x0[hm] + p.tap_dx[tap],
                            y0[hm] + p.tap_dy[tap], img[hm]);
              } else {
                chunk = kb - p.ntaps * p.cpb;
                tma_load_4d(sah, &maps.a[4], full_bar + stage, chunk * BK, x0[hm], y0[hm], img[hm]);
              }
            }
          }
          tma_load_3d(sb, &maps.b, full_bar + stage, b_k0 + kb * BK, b_n0, b_b);
          if (++stage == SM::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    const int iss = warp == 1 ? 0 : 1;
    if (lane == 0 && iss < p.issuers) {
      // =============================== MMA issuers ===============================
      // issuer `iss` takes the k-blocks kb with kb % issuers == iss; both walk the same (stage, phase) sequence.  The first
      // k-block of a tile (issuer 0) zeroes the accumulator: issuer 1 waits until it has retired (start_bar) before its
      // first accumulating MMA of that tile.
      const int NI = p.issuers;
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (p.causal_skip) {
          const int rem = tile % tiles_mn;
          const int tm = rem / p.tiles_n, tn = rem - tm * p.tiles_n;
          if (tn * BN > tm * GEMM_BM + GEMM_BM - 1) continue;
        }
        // MH == 1: the two accumulators alternate between tiles (double buffering); MH == 2: they are the two halves of the tile
        const int acc = MH == 1 ? (local & 1) : 0;
        const uint32_t acc_phase = MH == 1 ? ((local >> 1) & 1) : (local & 1);
        ++local;
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        if (MH == 2) mbar_wait(tempty_bar + 1, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        bool waited_start = iss == 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (NI == 1 || (kb & 1) == iss) {
            if (!waited_start) { mbar_wait(start_bar + acc, acc_phase); waited_start = true; }
            mbar_wait(full_bar + stage, phase);
            if (XF) mbar_wait(xf_bar + stage, phase);       // the transform warps have rewritten the A tile in place
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * SM::STAGE_BYTES);
            const uint64_t bdesc = umma_desc_sw128_kmajor(sa + MH * SM::A_BYTES);
#pragma unroll
            for (int hm = 0; hm < MH; ++hm) {
              const uint64_t adesc = umma_desc_sw128_kmajor(sa + hm * SM::A_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per k-block
                umma_ss<TF32>(tmem_d + (uint32_t)(hm * BN), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC,
                              (kb > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(empty_bar + stage);  // smem slot reusable once these MMAs retire
            if (NI == 2 && kb == 0) umma_commit(start_bar + acc);
          }
          if (++stage == SM::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + acc);      // this issuer's share of the accumulator(s) complete -> epilogue (NI arrivals)
        if (MH == 2) umma_commit(tfull_bar + 1);
      }
    }
  } else if (XF && warp >= 7) {
    static_assert(GEMM_XF_WARPS == 4 || GEMM_XF_WARPS == 8 || GEMM_XF_WARPS == 16, "transform warps: 4, 8 or 16");
    // =============================== operand-transform warps (fused GroupNorm + SiLU of the conv input) ===============================
    // 128 threads; thread (c = tt & 7, rbase = tt >> 3) owns the 16-byte chunk of channels [chunk0 + c * NCH, + NCH) of rows
    // rbase + 16 i: its per-(image, channel) coefficients stay in registers across the rows of a k-block and are prefetched one
    // k-block ahead.  A quarter warp covers the 8 (permuted) chunks of one 128-byte row: bank-conflict free.
    constexpr int NCH = XfChunk<T>::NCH;
    constexpr int XROWS = 128 / (4 * GEMM_XF_WARPS);       // rows per thread: 8 / 4 / 2
    constexpr int XSTEP = 4 * GEMM_XF_WARPS;               // row stride between them
    const int tt = threadIdx.x - GEMM_THREADS;
    const int c = tt & 7, rbase = tt >> 3;
    const int kb_taps = p.ntaps * p.cpb;
    const bool silu = p.xf_silu != 0;
    int rty[XROWS], rtx[XROWS];                           // tile-relative pixel of this thread's rows
#pragma unroll
    for (int i = 0; i < XROWS; ++i) { const int r = rbase + XSTEP * i; rty[i] = r / p.tw; rtx[i] = r - rty[i] * p.tw; }
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int rem = tile % tiles_mn;
      const int tm = rem / p.tiles_n;
      const int tiles_x = p.W / p.tw, tiles_img = tiles_x * (p.H / p.th);
      const int img = tm / tiles_img;
      const int r2 = tm - img * tiles_img;
      const int y0 = (r2 / tiles_x) * p.th, x0 = (r2 % tiles_x) * p.tw;
      const float* scp = p.xf_scale + (size_t)img * p.xf_cin + c * NCH;
      const float* shp = p.xf_shift + (size_t)img * p.xf_cin + c * NCH;
      const float cmul = silu ? 0.5f : 1.0f;               // silu(y) = h + h tanh(h), h = y / 2: fold the 1/2 into the affine map
      float sc[NCH], sh[NCH], nsc[NCH], nsh[NCH];
#pragma unroll
      for (int i = 0; i < NCH; i += 4) {
        *reinterpret_cast<float4*>(nsc + i) = __ldg(reinterpret_cast<const float4*>(scp + i));
        *reinterpret_cast<float4*>(nsh + i) = __ldg(reinterpret_cast<const float4*>(shp + i));
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool xf = kb < kb_taps;                       // k-blocks of the fused 1x1 shortcut source are left untouched
        int tap = 0;
        if (xf) {
#pragma unroll
          for (int i = 0; i < NCH; ++i) { sc[i] = nsc[i] * cmul; sh[i] = nsh[i] * cmul; }
          tap = kb / p.cpb;
          const int nkb = kb + 1;
          if (nkb < kb_taps) {                              // coefficients of the next k-block's channel chunk
            const int chunk0 = (nkb - (nkb / p.cpb) * p.cpb) * (BK);
#pragma unroll
            for (int i = 0; i < NCH; i += 4) {
              *reinterpret_cast<float4*>(nsc + i) = __ldg(reinterpret_cast<const float4*>(scp + chunk0 + i));
              *reinterpret_cast<float4*>(nsh + i) = __ldg(reinterpret_cast<const float4*>(shp + chunk0 + i));
            }
          }
        }
        mbar_wait(full_bar + stage, phase);
        if (xf) {
          uint8_t* sa = smem + stage * SM::STAGE_BYTES;
          const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
#pragma unroll
          for (int i = 0; i < XROWS; ++i) {
            const int r = rbase + XSTEP * i;
            const int yy = y0 + rty[i] + dy, xx = x0 + rtx[i] + dx;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {            // padding pixels were zero-filled by TMA and stay zero
              uint4* ptr = reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4));
              *ptr = XfChunk<T>::apply(*ptr, sc, sh, silu);
            }
          }
          fence_proxy_async();                              // generic-proxy stores -> visible to the tensor core's reads
        }
        mbar_arrive(xf_bar + stage);
        if (++stage == SM::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;        // row of the 128-row tile owned by this thread
    float* gn_acc = reinterpret_cast<float*>(smem + SM::GN_OFFSET) + q * 64;   // this warp's [32 groups][2]
    const bool gn_on = (p.gn_part != nullptr) && (p.mode == 1);
    const int gn_cpg = gn_on ? p.N / p.gn_groups : 1;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int bz = tile / tiles_mn;
      const int rem = tile - bz * tiles_mn;
      const int tm2 = rem / p.tiles_n, tn = rem - tm2 * p.tiles_n;
      const int n0 = tn * BN;
      if (p.causal_skip && n0 > tm2 * GEMM_BM + GEMM_BM - 1) continue;
      const int local_tile = local++;
      const int outer = bz / p.heads, h = bz - outer * p.heads;
      const int o_b = p.o_bsel == 0 ? 0 : (p.o_bsel == 1 ? outer : bz);
#pragma unroll 1
      for (int hm = 0; hm < MH; ++hm) {          // the M sub-tiles of this CTA tile (one unless MH == 2)
      const int tm = tm2 * MH + hm;
      const int m0 = tm * GEMM_BM;
      const int acc = MH == 1 ? (local_tile & 1) : hm;
      const uint32_t acc_phase = MH == 1 ? ((local_tile >> 1) & 1) : (local_tile & 1);

      // logical output row of this thread
      long long orow;
      bool row_ok;
      if (p.mode == 0) {
        orow = m0 + row;
        row_ok = (m0 + row) < p.M;
      } else {
        const int tiles_x = p.W / p.tw, tiles_img = tiles_x * (p.H / p.th);
        const int img = tm / tiles_img;
        const int r2 = tm - img * tiles_img;
        const int y = (r2 / tiles_x) * p.th + row / p.tw;
        const int x = (r2 % tiles_x) * p.tw + row % p.tw;
        orow = ((long long)img * p.H + y) * p.W + x;
        row_ok = orow < (long long)p.M;
      }
      const int ocol0 = n0 + h * p.o_nhead;  // column in the output matrix (before SwiGLU halving)

      if (gn_on) { gn_acc[lane] = 0.f; gn_acc[lane + 32] = 0.f; __syncwarp(); }
      const uint32_t taddr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
      // ---- fast path (warp-uniform conditions): no per-row bias, 16-byte aligned rows ----
      int c_done = 0;
      bool gn_fast_done = false;
      {
        const int es_o = p.out_dtype == DT_BF16 ? 2 : 4, es_r = p.res_dtype == DT_BF16 ? 2 : 4;
        const long long obase = (long long)o_b * p.out_bstride + ocol0, rbase_ = (long long)o_b * p.res_bstride + ocol0;
        // SwiGLU halves the column index of the output: (ocol0 + c) / 2 stays 8-element aligned for c % 16 == 0
        const long long obase_al = p.act == 2 ? ((long long)o_b * p.out_bstride + (ocol0 >> 1)) : obase;
        const int ncols_t = (p.N - n0) < BN ? (p.N - n0) : BN;
        // epilogue statistics ride on the fast path when the tile is made of whole 64-column windows and a window holds at
        // most 16 groups (>= 4 channels per group)
        const bool gn_fast_ok = !gn_on || (ncols_t % 64 == 0 && gn_cpg >= 4 && p.gn_groups <= 32 && p.act != 2 && p.out_dtype == DT_BF16);
        bool fast = gn_fast_ok && !(p.bias && p.bias_along_m) && (p.act != 2 || (ocol0 & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(p.out) + (uintptr_t)(obase_al * es_o)) & 15) == 0 && (p.ldo * es_o) % 16 == 0;
        if (p.residual) fast = fast && ((reinterpret_cast<uintptr_t>(p.residual) + (uintptr_t)(rbase_ * es_r)) & 15) == 0 && (p.ldr * es_r) % 16 == 0;
        if (fast) {
          const int ncols = (p.N - n0) < BN ? (p.N - n0) : BN;
          auto run = [&](auto gc_tag, auto bf_tag) {
            constexpr int GC = decltype(gc_tag)::value;
            constexpr bool RB = decltype(bf_tag)::value;
            const int ngroups = ncols / GC;
            if (ngroups == 0) return;
            const char* rrow = p.residual ? reinterpret_cast<const char*>(p.residual) + (size_t)(rbase_ + orow * p.ldr) * es_r : nullptr;
            const bool use_res = rrow != nullptr && row_ok;
            EpiRes<GC, RB> cur, nxt;
            if (use_res) cur.load(rrow);                      // before the accumulator is complete
            // ---- epilogue GroupNorm statistics of the values being written (GC == 64 only) ----
            float2* gn_rows = reinterpret_cast<float2*>(smem + SM::GNR_OFFSET);         // [128][17]
            float2* gn_blk = gn_rows + 128 * 17;                                        // [8][16]
            float2* gn_tile = gn_blk + 8 * 16;                                          // [32]
            const int et = (int)threadIdx.x - 64;                                       // 0..127 over the four epilogue warps
            if (gn_on) {
              if (et < 32) gn_tile[et] = make_float2(0.f, 0.f);
              asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(tfull_bar + acc, acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int gi = 0; gi < ngroups; ++gi) {
              const int c = gi * GC;
              uint32_t r[GC / 16][16];
#pragma unroll
              for (int i = 0; i < GC / 16; ++i) tmem_ld_32x32b_x16(taddr + (uint32_t)(c + 16 * i), r[i]);
              if (use_res && gi + 1 < ngroups) nxt.load(rrow + (size_t)(c + GC) * es_r);
              tmem_ld_wait();
              // statistics: running (sum, sum of squares) of the current GroupNorm group, flushed to this row's slot when the
              // column index crosses a group boundary (comparisons only, no division per element)
              const int g_first = gn_on ? (n0 + c) / gn_cpg : 0;
              int g_cur = g_first, g_next_col = (g_first + 1) * gn_cpg;
              float gs = 0.f, gq = 0.f;
              if (row_ok || gn_on) {
                const long long ooff = obase + orow * p.ldo + c;
#pragma unroll
                for (int i = 0; i < GC / 16; ++i) {
                  float v[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[i][j]) * p.alpha;
                  if (p.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + h * p.o_nhead + n0 + c + 16 * i);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float4 b4 = __ldg(bp + j); v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w; }
                  }
                  if (use_res) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += cur.get(16 * i + j);
                  }
                  if (p.act == 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
                  }
                  if (gn_on) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const int col = n0 + c + 16 * i + j;
                      if (col == g_next_col) {
                        gn_rows[row * 17 + (g_cur - g_first)] = make_float2(gs, gq);
                        gs = 0.f; gq = 0.f; ++g_cur; g_next_col += gn_cpg;
                      }
                      const float x = row_ok ? v[j] : 0.f;
                      gs += x; gq = fmaf(x, x, gq);
                    }
                    if (!row_ok) continue;
                  }
                  if (p.act == 2) {               // interleaved (gate, up) pairs -> 8 outputs at column (ocol0 + c + 16 i) / 2
                    float o8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o8[j] = silu_f(v[2 * j]) * v[2 * j + 1];
                    const long long so = obase_al + orow * p.ldo + ((c + 16 * i) >> 1);
                    if (p.out_dtype == DT_BF16) {
                      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + so) =
                          make_uint4(pack_bf16x2(o8[0], o8[1]), pack_bf16x2(o8[2], o8[3]), pack_bf16x2(o8[4], o8[5]), pack_bf16x2(o8[6], o8[7]));
                    } else {
                      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + so);
                      op[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
                      op[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
                    }
                    continue;
                  }
                  if (p.out_dtype == DT_BF16) {
                    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff + 16 * i);
                    op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                  } else {
                    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + ooff + 16 * i);
#pragma unroll
                    for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                  }
                }
              }
              if (gn_on) {
                // this window's slots: rows -> 16-row blocks -> window total, fixed orders (bit-reproducible), then into the tile's
                // per-group sums; three named barriers over the four epilogue warps per 64 columns
                gn_rows[row * 17 + (g_cur - g_first)] = make_float2(gs, gq);
                const int ns = g_cur - g_first + 1;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                {
                  const int slot = et & 15, rb = et >> 4;
                  if (slot < ns) {
                    float s = 0.f, qv = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) { const float2 t = gn_rows[(rb * 16 + rr) * 17 + slot]; s += t.x; qv += t.y; }
                    gn_blk[rb * 16 + slot] = make_float2(s, qv);
                  }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et < ns && g_first + et < 32) {
                  float s = 0.f, qv = 0.f;
#pragma unroll
                  for (int rb = 0; rb < 8; ++rb) { const float2 t = gn_blk[rb * 16 + et]; s += t.x; qv += t.y; }
                  float2 acc2 = gn_tile[g_first + et];
                  acc2.x += s; acc2.y += qv;
                  gn_tile[g_first + et] = acc2;
                }
              }
              if (use_res) cur = nxt;
            }
            if (gn_on) {
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (tm < p.tiles_m) {
                const int tiles_x = p.W / p.tw, tiles_img = tiles_x * (p.H / p.th);
                const int img = tm / tiles_img, t_in = tm - img * tiles_img;
                const int slabs = tiles_img * p.tiles_n * 4;
                const int sub = et >> 5, ln = et & 31;                    // slab layout keeps 4 sub-records per (tile, n-tile)
                float* dst = p.gn_part + (((size_t)img * slabs) + ((size_t)t_in * p.tiles_n + tn) * 4 + sub) * p.gn_groups * 2;
                if (ln < p.gn_groups) {
                  const float2 t = sub == 0 ? gn_tile[ln] : make_float2(0.f, 0.f);
                  dst[2 * ln] = t.x; dst[2 * ln + 1] = t.y;
                }
              }
              asm volatile("bar.sync 1, 128;" ::: "memory");
              gn_fast_done = true;
            }
            c_done = ngroups * GC;
          };
          // bias (along N) needs n0 + c 4-float aligned: n0 and c are multiples of 16
          if (p.res_dtype == DT_BF16 || !p.residual) run(std::integral_constant<int, 64>{}, std::true_type{});
          else run(std::integral_constant<int, 32>{}, std::false_type{});
        }
      }
      if (c_done == 0) { mbar_wait(tfull_bar + acc, acc_phase); tc_fence_after(); }
      const float bias_m = (p.bias && p.bias_along_m && row_ok) ? p.bias[(p.mode == 0 ? (m0 + row) : 0)] : 0.f;

#pragma unroll 1
      for (int c = c_done; c < BN; c += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(taddr + (uint32_t)c, r);
        tmem_ld_wait();
        if (!gn_on) { if (!row_ok) continue; }
        if (n0 + c >= p.N) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
        if (p.bias) {
          if (p.bias_along_m) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += bias_m;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += (n0 + c + j < p.N) ? __ldg(p.bias + h * p.o_nhead + n0 + c + j) : 0.f;
          }
        }
        if (p.act == 2) {
          // interleaved (gate, up) pairs -> 8 outputs
          const long long ooff = (long long)o_b * p.out_bstride + orow * p.ldo + ((ocol0 + c) >> 1);
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = silu_f(v[2 * j]) * v[2 * j + 1];
          const bool full = (n0 + c + 16 <= p.N);
          if (p.out_dtype == DT_BF16) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + ooff;
            if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
              uint4 w = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                   pack_bf16x2(o[6], o[7]));
              *reinterpret_cast<uint4*>(op) = w;
            } else {
              for (int j = 0; j < 8; ++j) if (n0 + c + 2 * j + 1 < p.N) op[j] = __float2bfloat16_rn(o[j]);
            }
          } else {
            float* op = reinterpret_cast<float*>(p.out) + ooff;
            for (int j = 0; j < 8; ++j) if (n0 + c + 2 * j + 1 < p.N) op[j] = o[j];
          }
          continue;
        }
        if (p.residual && row_ok) {
          const long long roff = (long long)o_b * p.res_bstride + orow * p.ldr + ocol0 + c;
          if (p.res_dtype == DT_BF16) {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + c + j < p.N) v[j] += __bfloat162float(rp[j]);
          } else {
            const float* rp = reinterpret_cast<const float*>(p.residual) + roff;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + c + j < p.N) v[j] += rp[j];
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
        }
        if (gn_on) {
          // statistics of the values being written (fp32, before rounding), reduced over the warp's 32 pixels; the
          // control flow below depends only on column indices, so it is warp-uniform.
          int g_prev = (n0 + c) / gn_cpg;
          float ss = 0.f, qq = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int g = (n0 + c + j) / gn_cpg;
            if (g != g_prev) {
#pragma unroll
              for (int off = 16; off >= 1; off >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, off); qq += __shfl_xor_sync(0xffffffffu, qq, off); }
              if (lane == 0) { gn_acc[2 * g_prev] += ss; gn_acc[2 * g_prev + 1] += qq; }
              ss = 0.f; qq = 0.f; g_prev = g;
            }
            const float x = (row_ok && n0 + c + j < p.N) ? v[j] : 0.f;
            ss += x; qq = fmaf(x, x, qq);
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, off); qq += __shfl_xor_sync(0xffffffffu, qq, off); }
          if (lane == 0 && g_prev < p.gn_groups) { gn_acc[2 * g_prev] += ss; gn_acc[2 * g_prev + 1] += qq; }
          if (!row_ok) continue;
        }
        const long long ooff = (long long)o_b * p.out_bstride + orow * p.ldo + ocol0 + c;
        const bool full = (n0 + c + 16 <= p.N);
        if (p.out_dtype == DT_BF16) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + ooff;
          if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
            uint4 w0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                  pack_bf16x2(v[6], v[7]));
            uint4 w1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                  pack_bf16x2(v[14], v[15]));
            reinterpret_cast<uint4*>(op)[0] = w0;
            reinterpret_cast<uint4*>(op)[1] = w1;
          } else {
            for (int j = 0; j < 16; ++j) if (n0 + c + j < p.N) op[j] = __float2bfloat16_rn(v[j]);
          }
        } else {
          float* op = reinterpret_cast<float*>(p.out) + ooff;
          if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            for (int j = 0; j < 16; ++j) if (n0 + c + j < p.N) op[j] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + acc);
      if (gn_on && !gn_fast_done && tm < p.tiles_m) {
        const int tiles_x = p.W / p.tw, tiles_img = tiles_x * (p.H / p.th);
        const int img = tm / tiles_img, t_in = tm - img * tiles_img;
        const int slabs = tiles_img * p.tiles_n * 4;
        float* dst = p.gn_part + (((size_t)img * slabs) + ((size_t)t_in * p.tiles_n + tn) * 4 + q) * p.gn_groups * 2;
        if (lane < p.gn_groups) { dst[2 * lane] = gn_acc[2 * lane]; dst[2 * lane + 1] = gn_acc[2 * lane + 1]; }
        __syncwarp();
      }
      }  // hm
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T, int BN, bool XF, int MH>
static int launch_xf(const GemmMaps& maps, const GemmParams& p, int num_sms, cudaStream_t stream) {
  using SM = GemmSmem<BN, MH>;
  static PerDeviceOnce attr_once;
  if (attr_once.pending()) {
    IVG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<T, BN, XF, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_once.mark();
  }
  long long total = (long long)((p.tiles_m + MH - 1) / MH) * p.tiles_n * p.batch;
  int grid = (int)(total < num_sms ? total : num_sms);
  if (grid < 1) return 0;
  IVG_CUDA(launch_k(gemm_tc_kernel<T, BN, XF, MH>, dim3(grid), dim3(XF ? GEMM_THREADS_XF : GEMM_THREADS), (size_t)SM::TOTAL, stream,
                    maps, p));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}
template <typename T, int BN>
static int launch_one(const GemmMaps& maps, const GemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.xf_scale != nullptr) {
    IVG_CHECK(p.mode == 1 && p.xf_shift != nullptr && p.xf_cin > 0, "gemm_tc: the operand transform belongs to stride-1 conv launches");
    return launch_xf<T, BN, true, 1>(maps, p, num_sms, stream);
  }
  if constexpr (BN == 256) {
    // 256 x 256 CTA tiles (two M sub-tiles on one weight tile) when there is enough work for every SM and nothing needs the
    // per-tile accumulator double buffering more than the operand traffic saved (no causal skipping, no epilogue statistics)
    if (p.mh2 && !p.causal_skip && p.gn_part == nullptr && (long long)(p.tiles_m / 2) * p.tiles_n * p.batch >= 2LL * num_sms) {
      GemmParams q = p;
      q.issuers = 1;       // (two issuers on these tiles faulted on the batched K = 128 attention-score launches; not understood)
      return launch_xf<T, BN, false, 2>(maps, q, num_sms, stream);
    }
  }
  return launch_xf<T, BN, false, 1>(maps, p, num_sms, stream);
}

// ---- optional per-launch event timing (bench.py roofline leg) -----------------------------------------------
struct ProfRec { cudaEvent_t e0, e1; double flops; int bucket; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;

void gemm_profile_enable(int on) { g_prof_on = on != 0; }

int gemm_profile_collect(int bucket, double* ms_total, double* flops_total, long long* launches) {
  double ms = 0.0, fl = 0.0; long long n = 0;
  std::vector<ProfRec> keep;
  for (auto& r : g_prof) {
    if (r.bucket != bucket) { keep.push_back(r); continue; }
    IVG_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    IVG_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t; fl += r.flops; ++n;
    g_prof_pool.push_back({r.e0, r.e1});
  }
  g_prof.swap(keep);
  *ms_total = ms; *flops_total = fl; *launches = n;
  return 0;
}

// generic bracket for launchers in other translation units (bucket 2 = decode megakernel): returns a record index or -1
int profile_begin(int bucket, double work, cudaStream_t stream) {
  if (!g_prof_on) return -1;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return -1;
  ProfRec r;
  if (!g_prof_pool.empty()) { r.e0 = g_prof_pool.back().first; r.e1 = g_prof_pool.back().second; g_prof_pool.pop_back(); }
  else { if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1; }
  r.flops = work;
  r.bucket = bucket;
  cudaEventRecord(r.e0, stream);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
void profile_end(int idx, cudaStream_t stream) {
  if (idx >= 0 && idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].e1, stream);
}

static int dispatch_inner(int dtype, int bn, const GemmMaps& maps, const GemmParams& p, int num_sms,
                          cudaStream_t stream) {
  if (dtype == DT_BF16) {
    switch (bn) {
      case 32: return launch_one<__nv_bfloat16, 32>(maps, p, num_sms, stream);
      case 64: return launch_one<__nv_bfloat16, 64>(maps, p, num_sms, stream);
      case 128: return launch_one<__nv_bfloat16, 128>(maps, p, num_sms, stream);
      case 256: return launch_one<__nv_bfloat16, 256>(maps, p, num_sms, stream);
    }
  } else if (dtype == DT_F32) {
    switch (bn) {
      case 32: return launch_one<float, 32>(maps, p, num_sms, stream);
      case 64: return launch_one<float, 64>(maps, p, num_sms, stream);
      case 128: return launch_one<float, 128>(maps, p, num_sms, stream);
      case 256: return launch_one<float, 256>(maps, p, num_sms, stream);
    }
  }
  set_error("gemm_tc: unsupported dtype=%d / BN=%d", dtype, bn);
  return 1;
}

int gemm_tc_dispatch(int dtype, int bn, const GemmMaps& maps, const GemmParams& p, int num_sms, cudaStream_t stream) {
  bool prof = g_prof_on;
  if (prof) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) prof = false;
  }
  if (!prof) return dispatch_inner(dtype, bn, maps, p, num_sms, stream);
  ProfRec r;
  if (!g_prof_pool.empty()) { r.e0 = g_prof_pool.back().first; r.e1 = g_prof_pool.back().second; g_prof_pool.pop_back(); }
  else { IVG_CUDA(cudaEventCreate(&r.e0)); IVG_CUDA(cudaEventCreate(&r.e1)); }
  const int BK = 128 / (dtype == DT_BF16 ? 2 : 4);
  double frac = 1.0;
  if (p.causal_skip) {  // only tiles on or below the diagonal are computed
    long long done = 0, all = (long long)p.tiles_m * p.tiles_n;
    for (int tm = 0; tm < p.tiles_m; ++tm)
      for (int tn = 0; tn < p.tiles_n; ++tn) done += (tn * bn <= tm * GEMM_BM + GEMM_BM - 1);
    frac = all ? (double)done / (double)all : 1.0;
  }
  r.flops = 2.0 * (double)p.M * (double)p.N * ((double)p.num_kb * BK) * (double)p.batch * frac;
  r.bucket = p.mode;
  IVG_CUDA(cudaEventRecord(r.e0, stream));
  int rc = dispatch_inner(dtype, bn, maps, p, num_sms, stream);
  IVG_CUDA(cudaEventRecord(r.e1, stream));
  g_prof.push_back(r);
  return rc;
}

}  // namespace ivg
