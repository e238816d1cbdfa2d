// Fused causal attention for the prefill / teacher-forced pass of the Llama-style transformer (head_dim 64, bf16):
//     O = softmax(Q K^T / sqrt(64) + causal mask) V        per (clip, head)
// replacing HF LlamaAttention's eager/SDPA attention behind reference inference/predict.py:64 (generate -> prefill) and
// train_gpt.py:792 (full-sequence forward).  Round 1 materialised the scores and probabilities in HBM (Q.K^T GEMM ->
// softmax kernel -> P.V GEMM: 821 MB + 410 MB per layer at B = 64); here S and P never leave the SM.
//
// One CTA per SM, persistent over (clip*head, 128-row query tile) work items, longest (most key tiles) first.
//   warp 0, lane 0 : TMA producer -- Q tile once per item, then a 2-stage ring of {K tile [128 keys x 64], V^T tile [64 x 128 keys]}
//   warp 1, lane 0 : tcgen05.mma issuer -- S_j = Q K_j^T into one of two 128-column TMEM buffers (S_{j+1} is issued before
//                    P_j is awaited, so the tensor core works on the next scores while the softmax warps work on these),
//                    then PV_j = P_j V_j into a 64-column TMEM buffer (NOT accumulated across j)
//   warps 2-5      : softmax + accumulation, one query row per thread: S_j from TMEM (tcgen05.ld), running max / sum
//                    (online softmax, base-2 exponentials), P_j as bf16 into shared memory in the 128B-swizzled K-major
//                    layout the tensor core reads, O (64 fp32 registers per thread) rescaled and += PV_j read from TMEM.
// Keeping O in registers instead of accumulating in TMEM avoids the read-modify-write of the accumulator when the
// running maximum moves; the price is one 64-column tcgen05.ld per key tile.
#include "common.cuh"

namespace ivg {

constexpr int FA_THREADS = 192;
constexpr int FA_BM = 128;            // query rows per work item
constexpr int FA_BN = 128;            // keys per tile
constexpr int FA_D = 64;              // head dim
constexpr int FA_Q_BYTES = FA_BM * 128;                  // 16 KB: [128 rows][64 bf16], one 128-byte swizzle atom per row
constexpr int FA_K_BYTES = FA_BN * 128;                  // 16 KB
constexpr int FA_V_BYTES = 2 * FA_D * 128;               // 16 KB: two k-blocks of [64 d-rows][64 keys]
constexpr int FA_STAGE_BYTES = FA_K_BYTES + FA_V_BYTES;  // 32 KB
constexpr int FA_STAGES = 2;
constexpr int FA_P_BYTES = 2 * FA_BM * 128;              // 32 KB: two k-blocks of [128 rows][64 keys]
constexpr int FA_BAR_OFFSET = FA_Q_BYTES + FA_STAGES * FA_STAGE_BYTES + FA_P_BYTES;
constexpr int FA_SMEM = FA_BAR_OFFSET + 256 + 1024;      // + barriers + alignment slack
constexpr int FA_TMEM_COLS = 512;                        // S0 [0,128) S1 [128,256) PV [256,320)

struct alignas(64) FlashMaps {
  CUtensorMap q;    // 3-D (64, Lq, BH)      box (64, 128, 1)
  CUtensorMap k;    // 3-D (64, Lk, BH)      box (64, 128, 1)
  CUtensorMap vt;   // 3-D (Lk, 64, BH)      box (64, 64, 1)
};

struct FlashParams {
  int BH, heads, Lq, Lk;
  int causal;         // 1: key j visible to query i iff j <= i + causal_off
  int causal_off;     // Lk - Lq for a chunk appended to an existing cache
  float scale_log2;   // softmax scale * log2(e)
  __nv_bfloat16* out; // [B * Lq, heads * 64]
  long long ldo;      // heads * 64
  float* lse;         // optional [BH, Lq]: log-sum-exp (natural log) of the scaled scores, for a backward pass
};

__device__ __forceinline__ int fa_num_kv(const FlashParams& p, int qt) {
  int nkv = (p.Lk + FA_BN - 1) / FA_BN;
  if (p.causal) {
    const int last_key = qt * FA_BM + FA_BM - 1 + p.causal_off;     // last key any row of the tile may see
    const int lim = last_key / FA_BN + 1;
    nkv = lim < nkv ? lim : nkv;
  }
  return nkv < 1 ? 1 : nkv;
}

__global__ void __launch_bounds__(FA_THREADS, 1) flash_attn_kernel(const __grid_constant__ FlashMaps maps, const FlashParams p) {
  extern __shared__ uint8_t fa_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sq = smem;
  uint8_t* skv = smem + FA_Q_BYTES;
  uint8_t* sp = skv + FA_STAGES * FA_STAGE_BYTES;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + FA_BAR_OFFSET);
  uint64_t* q_empty = q_full + 1;
  uint64_t* kv_full = q_full + 2;      // [2]
  uint64_t* kv_empty = q_full + 4;     // [2]
  uint64_t* s_full = q_full + 6;       // [2]
  uint64_t* p_full = q_full + 8;
  uint64_t* pv_full = q_full + 9;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(q_full + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.vt);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); mbar_init(s_full + s, 1); }
    mbar_init(p_full, 128); mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_holder, FA_TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int QT = (p.Lq + FA_BM - 1) / FA_BM;
  const int total = p.BH * QT;

  if (warp == 0) {
    if (lane == 0) {
      // =============================== TMA producer ===============================
      uint32_t item = 0, kvc = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++item) {
        const int qt = QT - 1 - t / p.BH, bh = t % p.BH;            // longest items first
        const int nkv = fa_num_kv(p, qt);
        mbar_wait(q_empty, (item & 1u) ^ 1u);
        mbar_expect_tx(q_full, FA_Q_BYTES);
        tma_load_3d(sq, &maps.q, q_full, 0, qt * FA_BM, bh);
        for (int j = 0; j < nkv; ++j, ++kvc) {
          const uint32_t st = kvc & 1u, ph = (kvc >> 1) & 1u;
          mbar_wait(kv_empty + st, ph ^ 1u);
          uint8_t* sk = skv + st * FA_STAGE_BYTES;
          mbar_expect_tx(kv_full + st, FA_STAGE_BYTES);
          tma_load_3d(sk, &maps.k, kv_full + st, 0, j * FA_BN, bh);
          tma_load_3d(sk + FA_K_BYTES, &maps.vt, kv_full + st, j * FA_BN, 0, bh);
          tma_load_3d(sk + FA_K_BYTES + FA_D * 128, &maps.vt, kv_full + st, j * FA_BN + 64, 0, bh);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // =============================== MMA issuer ===============================
      constexpr uint32_t IDESC_S = umma_idesc(1, FA_BM, FA_BN);
      constexpr uint32_t IDESC_PV = umma_idesc(1, FA_BM, FA_D);
      uint32_t item = 0, kvc = 0, sc = 0, pc = 0;
      const uint64_t qdesc = umma_desc_sw128_kmajor(smem_u32(sq));
      auto issue_s = [&](uint32_t kv_index, uint32_t s_index) {
        const uint32_t st = kv_index & 1u, ph = (kv_index >> 1) & 1u;
        mbar_wait(kv_full + st, ph);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128_kmajor(smem_u32(skv + st * FA_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss<false>(tmem_base + (s_index & 1u) * 128u, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), IDESC_S, k > 0 ? 1u : 0u);
        umma_commit(s_full + (s_index & 1u));
      };
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++item) {
        const int qt = QT - 1 - t / p.BH;
        const int nkv = fa_num_kv(p, qt);
        mbar_wait(q_full, item & 1u);
        tc_fence_after();
        issue_s(kvc, sc);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) issue_s(kvc + j + 1, sc + j + 1);       // next scores while the softmax warps work on S_j
          else umma_commit(q_empty);                               // all S MMAs of the item issued: Q may be replaced when they retire
          mbar_wait(p_full, pc & 1u);
          tc_fence_after();
          const uint32_t st = (kvc + j) & 1u;
          const uint32_t vbase = smem_u32(skv + st * FA_STAGE_BYTES + FA_K_BYTES);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t pdesc = umma_desc_sw128_kmajor(smem_u32(sp) + (uint32_t)(kb * FA_BM * 128));
            const uint64_t vdesc = umma_desc_sw128_kmajor(vbase + (uint32_t)(kb * FA_D * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<false>(tmem_base + 256u, pdesc + (uint64_t)(k * 2), vdesc + (uint64_t)(k * 2), IDESC_PV, (kb | k) ? 1u : 0u);
          }
          umma_commit(kv_empty + st);      // K_j / V_j (and P_j) consumed once these retire
          umma_commit(pv_full);
          ++pc;
        }
        kvc += (uint32_t)nkv;
        sc += (uint32_t)nkv;
      }
    }
  } else {
    // =============================== softmax / accumulation warps ===============================
    const int q4 = warp & 3;                         // TMEM lane quadrant of this warp
    const int r = q4 * 32 + lane;                    // row of the query tile owned by this thread
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    uint32_t sc = 0, pc = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int qt = QT - 1 - t / p.BH, bh = t % p.BH;
      const int nkv = fa_num_kv(p, qt);
      const int qi = qt * FA_BM + r;                 // query index of this thread
      const int vis = p.causal ? qi + p.causal_off : p.Lk - 1;     // last visible key
      float m = -INFINITY, l = 0.f;
      float o[FA_D];
#pragma unroll
      for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
      for (int j = 0; j < nkv; ++j, ++sc, ++pc) {
        mbar_wait(s_full + (sc & 1u), (sc >> 1) & 1u);
        tc_fence_after();
        const uint32_t saddr = tmem_base + lane_addr + (sc & 1u) * 128u;
        float s[FA_BN];
#pragma unroll
        for (int c = 0; c < FA_BN; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(saddr + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) s[c + i] = __uint_as_float(v[i]);
        }
        const int key0 = j * FA_BN;
        const int lastk = (vis < p.Lk - 1 ? vis : p.Lk - 1) - key0;   // tile-relative index of the last visible key
        float mx = m;
#pragma unroll
        for (int i = 0; i < FA_BN; ++i) {
          s[i] = i <= lastk ? s[i] * p.scale_log2 : -INFINITY;
          mx = fmaxf(mx, s[i]);
        }
        // rows with no visible key yet (cannot happen for causal_off >= 0 after the first tile) keep m = -inf: guard the exp
        const float mref = mx == -INFINITY ? 0.f : mx;
        const float alpha = exp2f(m - mref);
        float sum = 0.f;
        uint8_t* prow = sp + (size_t)r * 128;
#pragma unroll
        for (int c = 0; c < FA_BN; c += 8) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { e[i] = exp2f(s[c + i] - mref); sum += e[i]; }
          const uint4 w = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
          const int kb = c >> 6, chunk = (c & 63) >> 3;
          *reinterpret_cast<uint4*>(prow + (size_t)kb * (FA_BM * 128) + (size_t)((chunk ^ (r & 7)) << 4)) = w;
        }
        l = l * alpha + sum;
        m = mx;
#pragma unroll
        for (int i = 0; i < FA_D; ++i) o[i] *= alpha;
        fence_proxy_async();              // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        mbar_arrive(p_full);
        mbar_wait(pv_full, pc & 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < FA_D; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + 256u + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[c + i] += __uint_as_float(v[i]);
        }
        tc_fence_before();                // the next PV MMA (ordered after our next p_full arrival) may overwrite the buffer
      }
      if (qi < p.Lq) {
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        const int b = bh / p.heads, h = bh - b * p.heads;
        __nv_bfloat16* op = p.out + ((size_t)b * p.Lq + qi) * p.ldo + (size_t)h * FA_D;
#pragma unroll
        for (int c = 0; c < FA_D; c += 8) {
          const uint4 w = make_uint4(pack_bf16x2(o[c] * inv, o[c + 1] * inv), pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv),
                                     pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv), pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv));
          *reinterpret_cast<uint4*>(op + c) = w;
        }
        if (p.lse != nullptr) p.lse[(size_t)bh * p.Lq + qi] = (m + log2f(l)) * 0.6931471805599453f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, FA_TMEM_COLS); }
}

int flash_attn_launch(const void* q, const void* k, const void* vt, void* out, float* lse, int B, int heads, int Lq, int Lk,
                      long long q_bstride, long long k_bstride, long long vt_bstride, long long vt_ld, long long ldo, int causal,
                      float scale, int num_sms, cudaStream_t st) {
  IVG_CHECK(B >= 1 && heads >= 1 && Lq >= 1 && Lk >= 1, "flash_attn: bad shape B=%d heads=%d Lq=%d Lk=%d", B, heads, Lq, Lk);
  IVG_CHECK(!causal || Lk >= Lq, "flash_attn: causal attention needs Lk >= Lq (got %d < %d)", Lk, Lq);
  IVG_CHECK(vt_ld % 8 == 0 && ldo % 8 == 0, "flash_attn: V^T row pitch and output pitch must be multiples of 8 elements");
  FlashMaps maps;
  const uint64_t BH = (uint64_t)B * heads;
  {
    uint64_t dims[3] = {64, (uint64_t)Lq, BH};
    uint64_t str[2] = {128, (uint64_t)q_bstride * 2};
    uint32_t box[3] = {64, FA_BM, 1};
    if (make_tensor_map(&maps.q, DT_BF16, q, 3, dims, str, box, 1)) return 1;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)Lk, BH};
    uint64_t str[2] = {128, (uint64_t)k_bstride * 2};
    uint32_t box[3] = {64, FA_BN, 1};
    if (make_tensor_map(&maps.k, DT_BF16, k, 3, dims, str, box, 1)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)Lk, 64, BH};
    uint64_t str[2] = {(uint64_t)vt_ld * 2, (uint64_t)vt_bstride * 2};
    uint32_t box[3] = {64, FA_D, 1};
    if (make_tensor_map(&maps.vt, DT_BF16, vt, 3, dims, str, box, 1)) return 1;
  }
  FlashParams p;
  p.BH = (int)BH; p.heads = heads; p.Lq = Lq; p.Lk = Lk;
  p.causal = causal ? 1 : 0; p.causal_off = Lk - Lq;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.lse = lse;
  static PerDeviceOnce attr_once;
  if (attr_once.pending()) {
    IVG_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    attr_once.mark();
  }
  const long long total = (long long)BH * ((Lq + FA_BM - 1) / FA_BM);
  const int grid = (int)(total < num_sms ? total : num_sms);
  flash_attn_kernel<<<grid, FA_THREADS, FA_SMEM, st>>>(maps, p);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

}  // namespace ivg
