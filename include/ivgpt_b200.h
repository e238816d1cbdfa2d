/* ivgpt_b200 -- C ABI of the B200-native iVideoGPT next-frame-prediction hot path.
 *
 * The reference (thuml/iVideoGPT) has no native boundary of its own: its hot path is Python calling
 * torch / diffusers / transformers library kernels.  This header is the boundary a maintainer binds with
 * ctypes from the reference's Python modules; every entry point names the reference
 * call site(s) whose arithmetic it replaces.  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - every function returns 0 on success; non-zero on failure, with a thread-local message available from
 *     ivgpt_last_error().  Nothing ever calls exit()/abort().
 *   - all pointers are DEVICE pointers on the current CUDA device unless stated; `stream` is a cudaStream_t.
 *     Work is enqueued asynchronously on that stream; the library holds no global state besides per-kernel
 *     attribute caches, so calls are re-entrant per (device, stream).
 *   - activations are NHWC ([frames, H, W, C], C fastest), dtype IVGPT_F32 (fed to tensor cores as TF32) or
 *     IVGPT_BF16; accumulation is always fp32.
 *   - integer results (token ids) are int64, matching torch.argmin / generate outputs of the reference.
 */
#ifndef IVGPT_B200_H
#define IVGPT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define IVGPT_F32 0
#define IVGPT_BF16 1

#define IVGPT_ACT_NONE 0
#define IVGPT_ACT_SILU 1
#define IVGPT_ACT_SWIGLU 2 /* interleaved (gate, up) column pairs -> silu(gate)*up, N/2 output columns */

const char* ivgpt_last_error(void);
unsigned long long ivgpt_launch_count(void); /* kernels launched by this library since load */
int ivgpt_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Measurement hooks (bench.py): when enabled, every tensor-core GEMM (bucket 0) / conv (bucket 1) / decode-megakernel
 * (bucket 2) launch issued outside stream capture is bracketed by CUDA events on its own stream; collect() synchronises
 * them and returns the summed device time, the summed work (buckets 0/1: algorithmic FLOPs 2*M*N*K, causal tiles
 * excluded; bucket 2: decode steps -- bench.py turns steps into weight + KV bytes) and the launch count. */
int ivgpt_profile_enable(int on);
int ivgpt_profile_collect(int bucket, double* ms_total, double* flops_total, long long* launches);
int ivgpt_count_add(long long n); /* account for kernels replayed through a CUDA graph */

/* ---- VQ codebook lookup -------------------------------------------------------------------------
 * Replaces diffusers VectorQuantizer.forward's `argmin(cdist(z, E))`, called at
 * ivideogpt/vq_model/compressive_vq_model.py:199 (context codebook) and :202 (dynamics codebook).
 * z [N, D] fp32, codebook [K, D] fp32, idx [N] int64.  D must be 64.
 * Workspaces: enorm_ws [K] fp32, packed_ws [N] uint64.  Ties resolve to the lowest index. */
int ivgpt_vq_argmin(const float* z, const float* codebook, float* enorm_ws, unsigned long long* packed_ws,
                    long long* idx, int N, int K, int D, void* stream);
/* Summation order of the distance: 0 = one sequential FMA chain over d (FFMA kernel); 1 = even-d and odd-d chains added
 * at the end (packed fma.rn.f32x2 kernel).  oracle/vq_argmin_ref.c implements both; results are bit-exact per order. */
int ivgpt_vq_set_order(int order);
int ivgpt_vq_get_order(void);

/* ---- tensor-core GEMM --------------------------------------------------------------------------
 * out = epilogue(alpha * A . B^T): A [a_rows, a_cols] (lda), B [b_rows, b_cols] (ldb), K contiguous in both.
 * Replaces nn.Linear / 1x1 nn.Conv2d / torch.bmm call sites: Llama q,k,v,o,gate,up,down,lm_head
 * (transformers modeling_llama via inference/predict.py:64, train_gpt.py:792), quant_conv / post_quant_conv /
 * quant_linear / post_quant_linear (compressive_vq_model.py:188,196,241,245), nn.MultiheadAttention
 * projections and score/value products (conditional_vae.py:49), mid-block attention (diffusers Attention).
 * grid-z index bz in [0, batch): outer = bz / heads, h = bz % heads; *_bsel: 0 -> batch 0, 1 -> outer / *_bdiv,
 * 2 -> bz.  A columns start at a_kbase + h*a_khead; B columns at b_kbase + h*b_khead, B rows at n + h*b_nhead;
 * output columns at n + h*o_nhead. */
typedef struct ivgpt_gemm_desc {
  int dtype; /* operand dtype of A and B */
  int bn;    /* N tile: 32, 64, 128 or 256; 0 = choose */
  const void* a; long long lda, a_bstride; int a_rows, a_cols, a_batches;
  const void* b; long long ldb, b_bstride; int b_rows, b_cols, b_batches;
  int M, N, K;
  int batch, heads;
  int a_bsel, a_bdiv, b_bsel, b_bdiv, o_bsel;
  int a_kbase, a_khead, b_kbase, b_khead, b_nhead, o_nhead;
  int causal_skip;
  void* out; long long ldo, out_bstride; int out_dtype;
  const float* bias; int bias_along_m;
  const void* residual; long long ldr, res_bstride; int res_dtype;
  int act;
  float alpha;
} ivgpt_gemm_desc;
int ivgpt_gemm(const ivgpt_gemm_desc* d, void* stream);

/* ---- 3x3 convolution as implicit GEMM -----------------------------------------------------------
 * Replaces nn.Conv2d(k=3) inside diffusers ResnetBlock2D / Downsample2D / Upsample2D and the conv_in(64->C) /
 * conv_out(C->64) layers built at ivideogpt/vq_model/vae.py:86-137,236-294.
 * x NHWC [N, Hin, Win, Cin]; stride 1 => pad 1, stride 2 => pad (0,1,0,1) (diffusers downsample_padding=0).
 * w packed [Cout][9*Cin + C2], k index = (ky*3+kx)*Cin + c, then the C2 channels of the optional fused 1x1
 * source x2 (NHWC at the OUTPUT resolution) -- the ResnetBlock2D conv_shortcut.
 * out/residual NHWC [N, Hout, Wout, Cout]. */
typedef struct ivgpt_conv_desc {
  int dtype; int bn;
  const void* x; int N, Hin, Win, Cin; int stride;
  const void* w; int Cout;
  const void* x2; int C2;
  const float* bias;
  const void* residual; int res_dtype;
  int act;
  void* out; int out_dtype;
  /* optional fused GroupNorm statistics of the output (the nn.GroupNorm that consumes it): partial (sum, sum of
   * squares) per (frame, slab, group) written to gn_part [N][gn_slabs][gn_groups][2]; gn_slabs from ivgpt_conv3x3_plan,
   * finish with ivgpt_groupnorm_finalize.  NULL = off. */
  float* gn_part; int gn_groups;
  /* optional fused GroupNorm (+ SiLU) of the INPUT -- the `norm -> nonlinearity -> conv` of diffusers ResnetBlock2D (reached
   * from vae.py:133-137,236-294) and of the conv_norm_out -> conv_act -> conv_out tails (vae.py:188-193): the per-(frame,
   * channel) affine coefficients in_scale / in_shift [N][Cin] (ivgpt_groupnorm_coeff) are applied to every operand tile in
   * shared memory before the tensor core reads it; padding stays zero.  stride 1 only.  NULL = off. */
  const float* in_scale; const float* in_shift; int in_silu;
} ivgpt_conv_desc;
/* scale = rstd * gamma, shift = beta - mean * scale per (sample, channel) from GroupNorm statistics [samples][G][2] */
int ivgpt_groupnorm_coeff(const float* stats, const float* gamma, const float* beta, float* scale, float* shift, int samples,
                          int C, int G, void* stream);
int ivgpt_conv3x3(const ivgpt_conv_desc* d, void* stream);
/* tile width the launcher would pick for this problem (set d->bn to it to pin the choice) and the number of
 * GroupNorm partial slabs per frame it implies. */
int ivgpt_conv3x3_plan(const ivgpt_conv_desc* d, int* bn, int* gn_slabs);
/* stats [samples][G][2] = (mean, rstd) from partial sums part [samples][slabs][G][2]; count = elements per group. */
int ivgpt_groupnorm_finalize(const float* part, float* stats, int samples, int slabs, int G, double count, float eps,
                             void* stream);

/* ---- GroupNorm (nn.GroupNorm in ResnetBlock2D / conv_norm_out / CrossAttentionBlock) --------------
 * x [N, rows, C]; stats [N, G, 2] = (mean, rstd); part_ws [N * ceil(rows/64) * G * 2] fp32. */
int ivgpt_groupnorm_stats(int dtype, const void* x, float* part_ws, float* stats, int N, int rows, int C, int G,
                          float eps, void* stream);
/* y = (x-mean)*rstd*gamma+beta, optional SiLU, optional + pos[row % pos_rows][C] (conditional_vae.py:44-47).
 * coef_ws: fp32 workspace [2 * samples * C] for the per-(sample, channel) scale / shift. */
int ivgpt_groupnorm_apply(int dtype, const void* x, void* y, const float* stats, const float* gamma,
                          const float* beta, const float* pos, float* coef_ws, long long total_rows,
                          int rows_per_sample, int C, int G, int silu, int pos_rows, void* stream);

/* conv_in: 3x3 conv 3 -> Cout on NCHW fp32 pixels -> NHWC (vae.py:86,149).  w [Cout][27], b [Cout].
 * Frame n of the N processed frames is read from slot (n / frames_per_clip) * clip_frames + frame_offset +
 * n % frames_per_clip of the caller's [B, T, 3, H, W] clip tensor (the context / future split of
 * compressive_vq_model.py:170-171 without a copy). */
int ivgpt_conv_in(int dtype, const float* x_nchw, const float* w, const float* b, void* y, int N, int H, int W,
                  int Cout, int frames_per_clip, int clip_frames, int frame_offset, void* stream);
/* decoder head: GroupNorm -> SiLU -> 3x3 conv C -> 3, NHWC in, NCHW fp32 out (vae.py:292-294,363-369).
 * w [3][9][C] (tap-major, channel fastest), b [3].  Output frames are scattered into [B, T, 3, H, W] with the
 * same slot mapping as ivgpt_conv_in (the torch.cat of compressive_vq_model.py:274-277 without a copy). */
int ivgpt_conv_out3(int dtype, const void* x, const float* stats, const float* gamma, const float* beta,
                    const float* w, const float* b, float* y_nchw, int N, int H, int W, int C, int G,
                    int frames_per_clip, int clip_frames, int frame_offset, void* stream);
/* nearest 2x upsample, NHWC (diffusers Upsample2D). */
int ivgpt_upsample2x(int dtype, const void* x, void* y, int N, int H, int W, int C, void* stream);
/* (de)patchify, compressive_vq_model.py:192-195 / :247-250.  inverse=0: [F,R,R,C] -> [F*(R/P)^2, P*P*C]. */
int ivgpt_patchify(int dtype, const void* x, void* y, int F, int R, int C, int P, int inverse, void* stream);
int ivgpt_convert(int src_dtype, const void* x, int dst_dtype, void* y, long long n, void* stream);
/* The rest of diffusers VectorQuantizer.forward(beta, legacy=False) after the argmin (compressive_vq_model.py:297-301, the
 * training / evaluation forward :332-369): zq[n,:] = codebook[idx[n],:] in `dtype` (may be NULL) and
 * *loss = (beta + 1) * mean((zq - z)^2) -- the value of beta * mse(sg(zq), z) + mse(zq, sg(z)).  part_ws: fp32 [296] scratch. */
int ivgpt_vq_commit(int dtype, const float* z, const float* codebook, const long long* idx, void* zq, long long N, int D,
                    long long K, float beta, float* part_ws, float* loss, void* stream);

/* token (de)serialisation, compressive_vq_model.py:205-220 / :227-245 */
int ivgpt_tokens_serialise(const long long* idx_ctx, const long long* idx_dyn, long long* tokens, long long* labels,
                           int B, int t, int f, int cr, int dr, long long n_vq, long long n_dyn, void* stream);
/* bad_ctx (device int, may be NULL): set to 1 when a CONTEXT position holds an id outside [0, n_vq) -- the reference's
 * `self.quantize.embedding(...)` (:238) raises on such input; dynamics ids are clamped as the reference does (:236). */
int ivgpt_tokens_gather(int dtype, const long long* tokens, const float* cb_ctx, const float* cb_dyn, void* q_ctx,
                        void* q_dyn, int B, int t, int f, int cr, int dr, int D, long long n_vq, long long n_dyn,
                        int L, int* bad_ctx, void* stream);

/* ---- input pipeline -----------------------------------------------------------------------------------
 * Replaces inference/utils.py:12-16 NPZParser.preprocess (and ivideogpt/data/simple_dataloader.py:394,510):
 * `images / 255` followed by torchvision's resize to [out_h, out_w] (bilinear, antialias, align_corners=False).
 * frames: device uint8 (in_dtype 2) or fp32 (0), element (t, y, x, c) at t*stride_t + y*stride_y + x*stride_x + c*stride_c
 * (strides in elements: [T,H,W,C] episodes as stored in the .npz files, or [T,C,H,W]); out fp32 [T, C, out_h, out_w].
 * divisor = 255 for the reference's pipeline.  Down-scaling factors up to 19. */
int ivgpt_preprocess_resize(int in_dtype, const void* frames, long long stride_t, long long stride_y, long long stride_x,
                            long long stride_c, int T, int H, int W, int C, float* out, int out_h, int out_w, float divisor,
                            void* stream);

/* ---- Llama pieces (transformers LlamaForCausalLM as used by predict.py:64 / train_gpt.py:792) ----- */
int ivgpt_embed(const long long* ids, long long ids_stride, int L, const int* dpos, const float* table, float* x,
                long long M, int hidden, long long vocab, void* stream);
int ivgpt_add_rows(float* x, const float* e, long long n, void* stream);
int ivgpt_rmsnorm(int dtype, const float* x, const float* w, void* y, long long M, int hidden, float eps,
                  void* stream);
/* v_rows (optional, may be NULL): second copy of V laid out [B][heads][Lmax][64] like K -- what ivgpt_decode_mega's
 * attention phase streams; v_cache_t [B][heads][64][Lmax] is the K-major operand of the prefill P.V product. */
int ivgpt_rope_kv(int dtype, const void* qkv, void* q_out, void* k_cache, void* v_cache_t, void* v_rows, int B, int Lq,
                  int heads, int Lmax, int pos0, const int* dpos, const float* cos_tab, const float* sin_tab,
                  void* stream);
int ivgpt_softmax(int dtype, const float* S, void* P, long long rows, int Lq, int Lk, long long lds, long long ldp,
                  int causal, int causal_off, void* stream);
int ivgpt_decode_attn(int dtype, const void* q, const void* k_cache, const void* v_cache_t, void* out, int B,
                      int heads, int Lmax, int Lcur, const int* dpos, float scale, void* stream);
int ivgpt_argmax(const float* logits, long long ld, int rows, int V, long long* out, long long out_stride,
                 const int* dpos, void* stream);
int ivgpt_topk_sample(const float* logits, long long ld, int rows, int V, int k, float temperature,
                      unsigned long long seed, unsigned long long step, long long* out, long long out_stride,
                      const int* dpos, const unsigned long long* dseed /* added to seed when non-NULL */,
                      void* stream);
/* shifted CE of LlamaForCausalLM.forward(labels=...): logits [B,L,ld] fp32, labels [B,L] int64 (-100 ignored);
 * loss_rows / valid_ws [B*(L-1)] fp32 workspaces; loss_out[0] = mean over labelled positions, [1] = their count. */
int ivgpt_ce_loss(const float* logits, long long ld, int B, int L, int V, const long long* labels, float* loss_rows,
                  float* valid_ws, float* loss_out, void* stream);
int ivgpt_incr(int* p, int by, void* stream);
/* Forced separator slots of the action-conditioned rollout (ivideogpt/transformer/action_model.py:78-114): every
 * `period`-th position from slot0 holds `token` instead of a sampled one (:109-110) and the embedding fed there gets
 * slot_emb[b, i, :] = action_linear(a_i) added (:80-81).  Position is read from device memory (*dpos = position being
 * fed).  x [B, hidden] fp32, slot_emb [B, nslots, hidden] fp32, tokens [B, tok_stride] int64. */
int ivgpt_slot_embed_add(float* x, const float* slot_emb, const int* dpos, int B, int hidden, int slot0, int period,
                         int nslots, void* stream);
int ivgpt_slot_force(long long* tokens, long long tok_stride, const int* dpos, int B, int slot0, int period,
                     long long token, void* stream);
/* One decode step of attention for the newest token (HF generate's per-token forward with a KV cache): RoPE on q/k,
 * append k / v to the caches at position pos (= *dpos when dpos != NULL), attention over positions [0, pos].
 * qkv [B, 3*hidden], k_cache [B,heads,Lmax,64], v_cache_t [B,heads,64,Lmax], out [B, hidden]. */
int ivgpt_decode_attn_fused(int dtype, const void* qkv, void* k_cache, void* v_cache_t, void* out, int B, int heads,
                            int Lmax, int pos, const int* dpos, const float* cos_tab, const float* sin_tab,
                            float scale, void* stream);
/* Fused attention of the prefill / teacher-forced pass (bf16, head_dim 64): out[b, i, h*64:(h+1)*64] =
 * softmax(q_bh[i] . k_bh^T * scale (+ causal mask: key j visible iff j <= i + Lk - Lq)) . v_bh, scores and probabilities kept
 * on chip (TMEM / shared memory).  Replaces the attention inside HF LlamaAttention (predict.py:64 via generate, train_gpt.py:792).
 *   q  [B*heads][Lq][64] (batch pitch q_bstride elements), k [B*heads][Lk..][64] (pitch k_bstride),
 *   vt [B*heads][64][vt_ld >= Lk] (V transposed, pitch vt_bstride), out [B*Lq][ldo] bf16, lse optional fp32 [B*heads][Lq]. */
int ivgpt_flash_attn(const void* q, const void* k, const void* vt, void* out, float* lse, int B, int heads, int Lq, int Lk,
                     long long q_bstride, long long k_bstride, long long vt_bstride, long long vt_ld, long long ldo, int causal,
                     float scale, void* stream);
/* ---- persistent decode megakernel (bf16): `steps` consecutive decode steps of HF generate in ONE cooperative launch
 * (embed -> layers x {qkv, RoPE+append+attention, o, norm, gate/up+SwiGLU, down, norm} -> lm_head -> sample -> append).
 * Weights are streamed as whole 16-row K-slabs by 1-D bulk copies, so every matrix (wqkv [3h,h], wo [h,h], wgu
 * [2*inter,h] gate/up interleaved, wd [h,inter], lm_head [vocab,h]; bf16 row-major) is first re-laid once with
 * ivgpt_mega_pack_weight() into a buffer of ivgpt_mega_packed_elems(rows, cols) bf16.  The per-layer pointers live in
 * a device array of ivgpt_mega_layer_bytes()-sized records filled on the host with ivgpt_mega_fill_layer() and copied
 * to the device by the caller.  barrier/error must be zero before each launch. */
typedef struct ivgpt_mega_desc {
  int B, hidden, inter, heads, layers, vocab, Lmax, steps;
  int o_splits, d_splits; /* split-K of o-proj / down-proj; part holds max(o,d) x B x hidden fp32 */
  float eps;
  void *x, *xn, *qkv, *ao, *act, *part, *logits;
  long long ldl;
  void *kcache, *vcache;
  const float *embed, *norm_f, *cos_tab, *sin_tab;
  long long* tokens; long long tok_stride;
  int* dpos;
  int do_sample, topk; float inv_temp;
  const unsigned long long* dseed;
  unsigned int* barrier; int* error;
  const void* layers_dev; const void* lm_head_packed;
  long long* prof; /* optional device [24]: SM-cycle totals of CTA 0 per phase kind (norm, qkv, attention, o, gate/up,
                      down, lm_head, sample, barriers); NULL to disable */
  void* vrows;     /* bf16 [layers][B][heads][Lmax][64]: V cache in K's layout (filled by ivgpt_rope_kv), used and
                      appended to when attn_mode == 0 */
  void* attn_part; /* fp32 [8 * SMs][72] scratch and */
  void* attn_cnt;  /* uint32 [max(SMs, B * heads)] zeroed counters for attention items cut along the sequence (attn_mode 0:
                      the left-over items of a large batch in 4 parts; every item in floor(8 * SMs / (B * heads)) <= 8
                      parts when that is >= 2, e.g. B = 16) */
  int attn_mode;   /* attention phase: 0 = K/V streamed by 1-D bulk copies (TMA) into a shared-memory ring,
                      1 = register-staged loads */
  /* forced separator slots (action-conditioned rollout, action_model.py:78-114); slot_period == 0 disables.
   * Position q >= slot0 with (q - slot0) % slot_period == 0 is slot i = (q - slot0) / slot_period: its token is
   * slot_token (never sampled) and slot_emb[b, i, :] (fp32 [B, nslots, hidden], may be NULL) is added to its embedding. */
  int slot0, slot_period, nslots;
  long long slot_token;
  const float* slot_emb;
  int a_bulk;      /* must be 1: xn / ao / act are kept in global memory as the 128B-swizzled K-major shared-memory image of
                      the GEMM operand ([k/64][row < a_rows][chunk ^ (row & 7)][8]; a_rows = 64 if B <= 64 else 128 in
                      gemm_mode 0), so a phase loads its activation slab with ONE bulk copy; the three buffers hold a_rows
                      rows.  (0 was the row-major + cp.async path of earlier builds, removed.) */
  int mma_m64;     /* ignored: gemm_mode 0 uses the M=64 tcgen05.mma whenever B <= 64 */
  /* gemm_mode 1: weight-stationary GEMM phases -- the 64 weight rows of a work item are the MMA's M side, the batch its N
   * side.  The five weight matrices are then packed with ivgpt_mega_pack_weight64() (wgu with swiglu_pairs = 1), a_bulk
   * must be 1, a_rows (rows of the activation images, a multiple of 8 >= B) is the MMA's N, o_splits / d_splits may go
   * up to 12, and the qkv projection is split-K too: qkvp holds its fp32 partials [qkv_splits][B][3*hidden] (the
   * attention phase sums them), `qkv` is unused. */
  int gemm_mode, qkv_splits, a_rows;
  void* qkvp;
  int bn_wide;     /* gemm_mode 0: weight rows per work item of the gate/up and lm_head phases (multiple of 16, <= 64; 0 = 16).
                      wgu and lm_head must be packed with the same width (ivgpt_mega_pack_weight_bn); wider items cut those
                      phases from 3 / 7 rounds of 48 tcgen05.mma issues over the SMs to 1 / 3 */
  int bn_down;     /* gemm_mode 0: weight rows per work item of the down projection (multiple of 16, <= 64; 0 = 16); wd must be
                      packed with the same width.  Wider tiles x more K splits keep one round over the SMs with fewer
                      tcgen05.mma issues per CTA (32-row tiles x 6 splits: 32 instead of 64 for the 138 M model) */
  void* tile_cnt;  /* uint32 [128] zeroed: arrival counters of the o-proj / down-proj output tiles.  Needed when
                      ivgpt_mega_fused_norm() == 1 (gemm_mode 0): the add + RMSNorm phases are folded into their neighbours --
                      the last split-K item of a tile adds the partials to x and writes bf16(x) into `xn`; wqkv / wgu /
                      lm_head must then be packed from W (.) g (the RMSNorm weight of their input folded into the columns)
                      and the kernel multiplies the per-row 1/rms in the consumers' epilogues (63 instead of 87 device-wide
                      barriers per step of a 12-layer model). */
} ivgpt_mega_desc;
int ivgpt_mega_fused_norm(void);   /* compile-time property of gemm_mode 0, see tile_cnt */
/* GEMM / conv kernel, tcgen05.mma issuing warps: 1 (default, bit-reproducible) or 2 (opt-in: IVGPT_MMA_ISSUERS=2 or
 * ivgpt_set_mma_issuers(2)): two warps accumulate alternate k-blocks into one accumulator -- the full tensor rate in isolation,
 * no gain in the operand-bound kernel, fp32 summation order (last bits) not fixed.  ivgpt_set_deterministic(1) forces 1. */
int ivgpt_set_deterministic(int on);
int ivgpt_set_mma_issuers(int n);
/* 1: BN = 256 launches with enough tiles use 256 x 256 CTA tiles (two 128-row sub-tiles on one weight tile: a third less
 * operand traffic, but the accumulator is no longer double-buffered across tiles).  Measured slower; default 0
 * (IVGPT_GEMM_MH2=1 in the environment also enables it). */
int ivgpt_set_gemm_mh2(int on);
int ivgpt_mega_layer_bytes(void);
long long ivgpt_mega_packed_elems(int rows, int cols);
int ivgpt_mega_pack_weight(const void* w, void* out, int rows, int cols, void* stream);
int ivgpt_mega_fill_layer(void* host_layer, const void* wqkv_packed, const void* wo_packed, const void* wgu_packed,
                          const void* wd_packed, const float* n1, const float* n2);
long long ivgpt_mega_packed_elems_bn(int rows, int cols, int bn);   /* bn weight rows per tile instead of 16 */
int ivgpt_mega_pack_weight_bn(const void* w, void* out, int rows, int cols, int bn, void* stream);
long long ivgpt_mega_packed_elems64(int rows, int cols);
int ivgpt_mega_pack_weight64(const void* w, void* out, int rows, int cols, int swiglu_pairs, void* stream);
int ivgpt_decode_mega(const ivgpt_mega_desc* d, void* stream);
/* ---- training pieces: backward of LlamaForCausalLM.forward(labels) (train_gpt.py:792-798) and the AdamW update
 * (train_gpt.py:648-658,803).  All contractions of the backward pass are ivgpt_gemm calls on transposed operands. */
int ivgpt_transpose(int dtype, const void* in, void* out, int batch, int rows, int cols, long long ld_in,
                    long long ld_out, long long bs_in, long long bs_out, void* stream);
int ivgpt_swiglu(int dtype, int backward, const void* gu /* [M,2I] gate/up interleaved */, const void* dact, void* out,
                 long long n /* M*I */, void* stream);
int ivgpt_rmsnorm_bwd(int dtype, const float* x, const float* w, const void* dy, float* dres /* += */, float* dw_part,
                      float* dw, long long M, int hidden, float eps, void* stream);
int ivgpt_softmax_bwd(int dtype, const void* P, const float* dP, void* dS, long long rows, int Lq, int Lk, long long ld,
                      int causal, float scale, void* stream);
int ivgpt_rope_bwd(int dtype, const float* dq, const float* dk, const float* dv, void* dqkv, int B, int L, int heads,
                   const float* cos_tab, const float* sin_tab, void* stream);
int ivgpt_ce_bwd(int dtype, const float* logits, long long ld, int B, int L, int V, const long long* labels,
                 const float* count, float gscale, void* dlogits, long long ldd, void* stream);
int ivgpt_embed_bwd(const long long* ids, const float* dx, float* dE, long long M, int hidden, long long vocab,
                    void* stream);
int ivgpt_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, float gscale, void* stream);
int ivgpt_add_to_f32(int dtype, float* y, const void* x, long long n, void* stream);
/* Attention dropout of the training step (HF LlamaAttention, `attention_dropout` of the Llama config;
 * scripts/pretrain/oxe-64-act-free.sh:31 uses 0.1): y[i] = keep(seed, i) ? x[i] / (1 - p) : 0, counter-based so that the
 * backward pass regenerates the forward mask from the seed.  y may alias x. */
int ivgpt_dropout(int dtype, const void* x, void* y, long long n, float p, unsigned long long seed, void* stream);
/* ---- tokenizer training: backward of CompressiveVQModel.forward (reference train_tokenizer.py:734 through
 * compressive_vq_model.py:332-369, vae.py:141-195,298-371, conditional_vae.py:38-55,108-132,186-212).  The contractions (conv
 * dgrad = ivgpt_conv3x3 on flipped weights, conv wgrad = dY^T x im2col(X)^T, Linear / attention products) are ivgpt_gemm /
 * ivgpt_conv3x3 calls; these are the fp32 NHWC kernels around them. */
int ivgpt_colsum(const float* x, long long M, int C, long long ld, float* part_ws /* part_rows*C */, int part_rows, float* out,
                 int accumulate, void* stream);
int ivgpt_groupnorm_bwd_chunks(int samples, int rows);   /* workspace sizing: ws = samples*(chunks+1)*C*2 + samples*G*2 floats */
int ivgpt_groupnorm_bwd(const float* x, const float* dy, const float* stats, const float* gamma, const float* beta, int silu,
                        int samples, int rows, int C, int G, float* ws, float* dx, int dx_accumulate, float* dgamma,
                        float* dbeta, int param_accumulate, void* stream);
int ivgpt_im2col3x3_t(const float* x, float* colT /* [k_rows][N*Ho*Wo] */, int N, int H, int W, int C, int stride, int k_rows,
                      void* stream);
/* out[c][n*img_stride + (y+1)*Wp + (x+1) - shift] = x[n][y][x][c] (out pre-zeroed by the caller, row pitch ld_out; Wp >= W+2, a
 * multiple of 4; shift in {-1,0,1}): the K-major operands of the 3x3 stride-1 weight gradient -- tap (a, b) reads the copy
 * written with shift = b-1 at K offset (a-1)*Wp (TMA boxes must start on 16-byte boundaries of the innermost dimension). */
int ivgpt_transpose_pad(const float* x, float* out, int N, int H, int W, int C, int Wp, int shift, long long img_stride,
                        long long ld_out, void* stream);
int ivgpt_zero_insert2x(const float* dy, float* out, int N, int h, int w, int C, void* stream);
int ivgpt_upsample2x_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream);
int ivgpt_silu(const float* x, const float* dy /* null: forward */, float* out, long long n, void* stream);
int ivgpt_axpby(const float* x, const float* y /* may be null */, float* out, float a, float b, long long n, void* stream);
int ivgpt_reduce_mid(const float* x, float* out, long long outer, int mid, long long inner, int accumulate, void* stream);
int ivgpt_nchw_to_nhwc(const float* x, float* y, long long N, int Cs, int Cd, long long HW, void* stream);
int ivgpt_vq_bwd(const float* z, const float* zq, const float* dout /* may be null */, const float* gloss /* device scalar or null */,
                 float beta, long long n, float* dz, float* de_rows, void* stream);
/* Programmatic dependent launch for the kernels of the decode step (prologue overlap inside CUDA graphs). */
int ivgpt_set_pdl(int on);

#ifdef __cplusplus
}
#endif
#endif /* IVGPT_B200_H */
