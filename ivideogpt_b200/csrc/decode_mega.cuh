// Parameter blocks of the persistent decode megakernel, shared by decode_mega.cu (device) and capi.cu (host).
#pragma once
#include "common.cuh"

namespace ivg {

constexpr int MEGA_THREADS = 256;   // 256: one warp per attention item (fastest measured); 512: warp pairs
constexpr int MEGA_BN = 16;                 // weight rows per GEMM work item
// gemm_mode 0: MEGA_NISSUE warps issue the tcgen05.mma instructions of a work item (lane 0 of warps 0..MEGA_NISSUE-1), each
// into its OWN TMEM accumulator; k-step t goes to issuer / accumulator t % MEGA_NISSUE and the epilogue adds the accumulators
// in a fixed order.  Why: one thread issuing back to back is limited to one tcgen05.mma per ~112 cycles whatever the shape
// (M <= 128, N <= 128); 4 issuing warps sustain ~40 cycles per instruction (tools/probes/mma_probe.cu,
// profiles/r02/mma_issue_and_ffma_probe.json).  A step of the 138 M model issues ~1800 MMAs per CTA.
#ifndef IVG_MEGA_NISSUE
#define IVG_MEGA_NISSUE 4
#endif
constexpr int MEGA_NISSUE = IVG_MEGA_NISSUE;   // 1, 2 or 4
// gemm_mode 0, fused norms (default): the two add + RMSNorm phases of a layer (and their device-wide barriers) do not exist.
//   * the residual add of a split-K projection (o, down) is done by the LAST work item of an output tile to arrive (monotonic
//     counter per tile): it sums the fp32 partials in split order (fixed -> bit-reproducible), updates x and writes the bf16
//     image of the UN-normalised row;
//   * RMSNorm is linear up to the per-row scale: xn @ W^T = rstd * ((x (.) g) @ W^T) = rstd * (x @ (W (.) g)^T).  The norm
//     weight g is folded into the columns of the consuming matrices when they are packed (wqkv, wgu, lm_head), the consumer's
//     A operand is bf16(x), and every CTA computes rstd = rsqrt(mean(x^2) + eps) of the rows from the activation slab it has
//     in shared memory anyway while the tensor core works, multiplying it in the epilogue.
// 63 instead of 87 device-wide barriers per decode step of the 12-layer model.
// MEASURED (profiles/r02/mega_fused_norm_ab.txt): not a win as built -- see DESIGN.md; kept as a compile-time variant, OFF.
#ifndef IVG_MEGA_FUSE_NORM
#define IVG_MEGA_FUSE_NORM 0
#endif
constexpr bool MEGA_FUSE_NORM = IVG_MEGA_FUSE_NORM != 0;
constexpr int MEGA_NACC = MEGA_NISSUE;         // one accumulator (64 TMEM columns) per issuing warp
constexpr int MEGA_MAXK = 1024;             // K handled by one work item (hidden or inter/3 ... all <= 1024)
constexpr int MEGA_A_BYTES = 128 * 1024;    // 64 rows x 1024 k x 2 B, or 128 rows x 512 ...; see a_rows below
constexpr int MEGA_B_BYTES = MEGA_BN * MEGA_MAXK * 2;   // 32 KB per slab
constexpr int MEGA_BAR_BYTES = 1024;           // mbarriers: 2 weight slabs, MMA done, TMEM holder, 8 x 8 attention-ring slots
constexpr int MEGA_SC_BYTES = 30 * 1024;       // attention scratch: per warp scores[Lmax + 8] + q[64] fp32
constexpr int MEGA_RING_SLOT = 4096;           // attention K/V ring slot (32 K rows, or up to 8 V^T rows)
constexpr int MEGA_SMEM = MEGA_A_BYTES + 2 * MEGA_B_BYTES + 1024 /*align*/ + MEGA_BAR_BYTES + MEGA_SC_BYTES;
constexpr int MEGA_MAX_SPLITS = 12;
// weight-stationary GEMM mode (gemm_mode 1): the WEIGHTS are the M side of the MMA (64 rows per work item), the batch is N
constexpr int MEGA_WM = 64;                 // weight rows per work item
constexpr int MEGA_W_CHUNK = 128;           // K per weight slab: 64 rows x 128 x 2 B = 16 KB, up to 4 in flight

struct MegaLayer {
  const __nv_bfloat16 *wqkv, *wo, *wgu, *wd;   // B operands packed by mega_pack_weight (16-row swizzled slab images)
  const float* n1;
  const float* n2;
};

struct MegaParams {
  int B, hidden, inter, heads, layers, vocab, Lmax, steps;
  int o_splits, d_splits;   // split-K factors of the o-proj (K = hidden) and down-proj (K = inter) phases
  float eps;
  float* x;                 // [B, hidden] fp32 residual stream
  __nv_bfloat16* xn;        // [B, hidden]
  __nv_bfloat16* qkv;       // [B, 3*hidden]
  __nv_bfloat16* ao;        // [B, hidden]
  __nv_bfloat16* act;       // [B, inter]
  float* part;              // [max(o_splits, d_splits)][B][hidden]
  float* logits;            // [B, ldl]
  long long ldl;
  __nv_bfloat16* kcache;    // [layers][B][heads][Lmax][64]
  __nv_bfloat16* vcache;    // [layers][B][heads][64][Lmax]
  __nv_bfloat16* vrows;     // [layers][B][heads][Lmax][64] (attn_mode 0)
  const float* embed;       // [vocab, hidden] fp32
  const float* norm_f;      // final RMSNorm weight
  const float* cos_tab;     // [max_pos, 32]
  const float* sin_tab;
  long long* tokens;        // [B, tok_stride]
  long long tok_stride;
  int* dpos;                // in: position of the token fed at step 0; out: advanced by `steps`
  int do_sample, topk;
  float inv_temp;
  const unsigned long long* dseed;
  unsigned int* barrier;    // zero-initialised
  int* error;               // zero-initialised; 1 = barrier timeout
  const MegaLayer* lw;      // device array [layers]
  const __nv_bfloat16* lm_head; // packed like the layer weights, rows padded to a multiple of 16
  long long* prof;              // optional [24] cycle counters filled by CTA 0 (phase breakdown; 9..13 attention, 14..17 GEMM item), may be null
  float* attn_part;             // [SMs][4][72] flash-decoding partials of the items cut along the sequence (attn_mode 0)
  unsigned int* attn_cnt;       // [SMs] zero-initialised arrival counters of those items
  unsigned int* tile_cnt;       // [2][64] zero-initialised arrival counters of the o-proj / down-proj output tiles (fused norms)
  int attn_mode;                // 0: TMA bulk-copy ring (default), 1: register-staged loads (round-1 v2 path)
  int bn_down;                  // gemm_mode 0: weight rows per work item of the down projection (16..64); wd is packed with it
  int bn_wide;                  // gemm_mode 0: weight rows per work item of the gate/up and lm_head phases (16..64); wgu and lm_head are packed with it
  int gemm_mode;                // 0: activations are the MMA's M side, 16 weight rows per item; 1: weight-stationary (see decode_mega.cu)
  int qkv_splits;               // gemm_mode 1: split-K of the qkv projection (partials in qkvp, summed by the attention prologue)
  float* qkvp;                  // gemm_mode 1: [qkv_splits][B][3*hidden] fp32
  int a_rows;                   // rows per k-block of the activation images / shared-memory activation slab
  int a_bulk;                   // 1: xn / ao / act are swizzled shared-memory images in global memory, loaded by one bulk copy
  int mma_m64;                  // 1: M = 64 UMMA in the GEMM phases when B <= 64 (0: M = 128 with the upper rows unused)
  // forced separator slots of the action-conditioned rollout (action_model.py:78-114); slot_period == 0 disables
  int slot0, slot_period, nslots;
  long long slot_token;
  const float* slot_emb;        // [B, nslots, hidden] fp32 or null
};

int decode_mega_launch(const MegaParams& p, int num_sms, cudaStream_t st);
int mega_fused_norm();        // 1 when gemm_mode 0 of this build folds the norms (the host packs g-folded weights)
int mega_pack_weight_launch(const void* w, void* out, int rows, int cols, int bn, cudaStream_t st);
int mega_pack_weight64_launch(const void* w, void* out, int rows, int cols, int swiglu_pairs, cudaStream_t st);

}  // namespace ivg
