// Non-GEMM kernels of the Llama-style autoregressive transformer (transformers LlamaForCausalLM, driven from
// reference inference/predict.py:64 (generate) and train_gpt.py:792 (forward with labels)):
// embedding gather, RMSNorm, RoPE + KV-cache append, row softmax, single-token decode attention over the
// cache, greedy / top-k sampling, shifted cross-entropy.  Residual stream is fp32; GEMM operands are T.
#include "common.cuh"

namespace ivg {

// ---------------------------------------------------------------------------------------------
// embedding gather: x[m, :] = E[ids[m], :] (+ extra[m, :] if given, e.g. action embeddings at sdf slots,
// action_model.py:80-81).  fp32 table -> fp32 residual stream.
// ---------------------------------------------------------------------------------------------
__global__ void embed_kernel(const long long* __restrict__ ids, long long ids_stride, int L,
                             const int* __restrict__ dpos, const float* __restrict__ table,
                             float* __restrict__ x, long long M, int Hd, long long vocab) {
  pdl_wait();
  pdl_launch_dependents();
  const int hv = Hd / 4;
  const long long total = M * hv;
  const int off = dpos ? *dpos : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / hv;
    const int c = (int)(i - m * hv);
    long long id = ids[(m / L) * ids_stride + (m % L) + off];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    reinterpret_cast<float4*>(x)[i] = __ldg(reinterpret_cast<const float4*>(table + id * Hd) + c);
  }
}

int embed_launch(const long long* ids, long long ids_stride, int L, const int* dpos, const float* table, float* x,
                 long long M, int Hd, long long vocab, cudaStream_t st) {
  IVG_CHECK(Hd % 4 == 0, "embed: hidden %% 4 != 0");
  if (M == 0) return 0;
  long long work = M * (Hd / 4);
  int blocks = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
  IVG_CUDA(launch_k(embed_kernel, dim3(blocks), dim3(256), 0, st, ids, ids_stride, L, dpos, table, x, M, Hd, vocab));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// RMSNorm (LlamaRMSNorm): y = w * (x * rsqrt(mean(x^2) + eps)), fp32 math, output T.  One warp per row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                               long long M, int Hd, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * Hd);
  const int nv = Hd / 4;
  float ss = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float4 v = xr[i];
    ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float r = rsqrtf(ss / (float)Hd + eps);
  T* yr = y + row * Hd;
  for (int i = lane; i < nv; i += 32) {
    float4 v = xr[i];
    float4 g = __ldg(reinterpret_cast<const float4*>(w) + i);
    yr[4 * i + 0] = from_f32<T>(g.x * (v.x * r));
    yr[4 * i + 1] = from_f32<T>(g.y * (v.y * r));
    yr[4 * i + 2] = from_f32<T>(g.z * (v.z * r));
    yr[4 * i + 3] = from_f32<T>(g.w * (v.w * r));
  }
}

int rmsnorm_launch(int dtype, const float* x, const float* w, void* y, long long M, int Hd, float eps,
                   cudaStream_t st) {
  IVG_CHECK(Hd % 4 == 0, "rmsnorm: hidden %% 4 != 0");
  if (M == 0) return 0;
  int blocks = (int)((M + 7) / 8);
  if (dtype == DT_BF16)
    IVG_CUDA(launch_k(rmsnorm_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, x, w, (__nv_bfloat16*)y, M, Hd, eps));
  else
    IVG_CUDA(launch_k(rmsnorm_kernel<float>, dim3(blocks), dim3(256), 0, st, x, w, (float*)y, M, Hd, eps));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// RoPE (half-split rotate_half; cos/sin tables [Lmax,32] precomputed in fp32 exactly as LlamaRotaryEmbedding does) applied to q and k of a fused qkv buffer
// [B*Lq, 3*h] (T); q -> qout [B, heads, Lq, 64]; k -> kcache [B, heads, Lmax, 64] at pos0 + l;
// v -> vcache transposed [B, heads, 64, Lmax].  head_dim is fixed at 64 (both Llama configs).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rope_kv_kernel(const T* __restrict__ qkv, T* __restrict__ qout, T* __restrict__ kcache,
                               T* __restrict__ vcache, T* __restrict__ vrows, int B, int Lq, int heads, int Lmax,
                               int pos0, const int* __restrict__ dpos, const float* __restrict__ cos_tab,
                               const float* __restrict__ sin_tab) {
  // thread = (b, l, head, i in 0..31)
  if (dpos) pos0 += *dpos;
  const long long total = (long long)B * Lq * heads * 32;
  const int Hd = heads * 64;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx & 31);
    long long r = idx >> 5;
    const int hh = (int)(r % heads); r /= heads;
    const int l = (int)(r % Lq);
    const int b = (int)(r / Lq);
    const int pos = pos0 + l;
    const float cs = __ldg(cos_tab + (size_t)pos * 32 + i), sn = __ldg(sin_tab + (size_t)pos * 32 + i);
    const T* row = qkv + ((size_t)b * Lq + l) * (3 * Hd);
    const float q0 = to_f32(row[hh * 64 + i]), q1 = to_f32(row[hh * 64 + i + 32]);
    const float k0 = to_f32(row[Hd + hh * 64 + i]), k1 = to_f32(row[Hd + hh * 64 + i + 32]);
    T* qo = qout + (((size_t)b * heads + hh) * Lq + l) * 64;
    qo[i] = from_f32<T>(q0 * cs - q1 * sn);
    qo[i + 32] = from_f32<T>(q1 * cs + q0 * sn);
    T* ko = kcache + (((size_t)b * heads + hh) * Lmax + pos) * 64;
    ko[i] = from_f32<T>(k0 * cs - k1 * sn);
    ko[i + 32] = from_f32<T>(k1 * cs + k0 * sn);
    T* vo = vcache + ((size_t)b * heads + hh) * 64 * Lmax + pos;
    vo[(size_t)i * Lmax] = row[2 * Hd + hh * 64 + i];
    vo[(size_t)(i + 32) * Lmax] = row[2 * Hd + hh * 64 + i + 32];
    if (vrows) {      // second copy, [pos][64] like K: the layout the decode megakernel streams (decode_mega.cu)
      T* vr = vrows + (((size_t)b * heads + hh) * Lmax + pos) * 64;
      vr[i] = row[2 * Hd + hh * 64 + i];
      vr[i + 32] = row[2 * Hd + hh * 64 + i + 32];
    }
  }
}

// Prefill-sized variant (round 2): one CTA per (clip, head, 64 positions).  The scalar kernel above moves 2 bytes per access and
// scatters V^T with a stride of Lmax elements: 245 us per layer at B = 64 (350 MB: 1.4 TB/s).  Here every global access is an
// 8-byte vector along the head dimension (q, K, the row-major V copy) and V^T goes through a shared-memory tile so that its
// rows (64 consecutive positions of one dimension) are written as 4-byte pairs, contiguous per warp.  bf16 only (the dtype of
// the benchmarked prefill); same arithmetic, same rounding as the scalar kernel.
__global__ void __launch_bounds__(256) rope_kv_tiled_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ qout,
                                                            __nv_bfloat16* __restrict__ kcache, __nv_bfloat16* __restrict__ vcache,
                                                            __nv_bfloat16* __restrict__ vrows, int Lq, int heads, int Lmax, int pos0,
                                                            const float* __restrict__ cos_tab, const float* __restrict__ sin_tab) {
  __shared__ __nv_bfloat16 vt[64][66];                      // [d][l], padded: conflict-free transposed reads
  const int l0 = blockIdx.x * 64, hh = blockIdx.y, b = blockIdx.z;
  const int Hd = heads * 64;
  const int t = threadIdx.x;
  const int iq = (t & 7) * 4;                               // dims iq..iq+3 and iq+32..iq+35
  // ---- q, k: rotate; v: copy + stash ----
  for (int l = t >> 3; l < 64; l += 32) {
    const int lq = l0 + l;
    if (lq >= Lq) break;
    const int pos = pos0 + lq;
    const __nv_bfloat16* row = qkv + ((size_t)b * Lq + lq) * (3 * Hd) + hh * 64;
    const float4 cs = __ldg(reinterpret_cast<const float4*>(cos_tab + (size_t)pos * 32 + iq));
    const float4 sn = __ldg(reinterpret_cast<const float4*>(sin_tab + (size_t)pos * 32 + iq));
    const float c4[4] = {cs.x, cs.y, cs.z, cs.w}, s4[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
    for (int which = 0; which < 2; ++which) {               // 0: q, 1: k
      const __nv_bfloat16* src = row + which * Hd;
      const uint2 lo = *reinterpret_cast<const uint2*>(src + iq), hi = *reinterpret_cast<const uint2*>(src + iq + 32);
      const __nv_bfloat16* lo_h = reinterpret_cast<const __nv_bfloat16*>(&lo);
      const __nv_bfloat16* hi_h = reinterpret_cast<const __nv_bfloat16*>(&hi);
      __nv_bfloat16 olo[4], ohi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x0 = __bfloat162float(lo_h[j]), x1 = __bfloat162float(hi_h[j]);
        olo[j] = __float2bfloat16_rn(x0 * c4[j] - x1 * s4[j]);
        ohi[j] = __float2bfloat16_rn(x1 * c4[j] + x0 * s4[j]);
      }
      __nv_bfloat16* dst = which == 0 ? qout + (((size_t)b * heads + hh) * Lq + lq) * 64
                                      : kcache + (((size_t)b * heads + hh) * Lmax + pos) * 64;
      *reinterpret_cast<uint2*>(dst + iq) = *reinterpret_cast<const uint2*>(olo);
      *reinterpret_cast<uint2*>(dst + iq + 32) = *reinterpret_cast<const uint2*>(ohi);
    }
    const uint2 vlo = *reinterpret_cast<const uint2*>(row + 2 * Hd + iq), vhi = *reinterpret_cast<const uint2*>(row + 2 * Hd + iq + 32);
    if (vrows) {
      __nv_bfloat16* vr = vrows + (((size_t)b * heads + hh) * Lmax + pos) * 64;
      *reinterpret_cast<uint2*>(vr + iq) = vlo;
      *reinterpret_cast<uint2*>(vr + iq + 32) = vhi;
    }
    const __nv_bfloat16* vl = reinterpret_cast<const __nv_bfloat16*>(&vlo);
    const __nv_bfloat16* vh = reinterpret_cast<const __nv_bfloat16*>(&vhi);
#pragma unroll
    for (int j = 0; j < 4; ++j) { vt[iq + j][l] = vl[j]; vt[iq + 32 + j][l] = vh[j]; }
  }
  __syncthreads();
  // ---- V^T rows: dimension d, positions l0 .. l0+63 (pairs of positions per lane) ----
  const int nl = (Lq - l0) < 64 ? (Lq - l0) : 64;
  for (int d = t >> 5; d < 64; d += 8) {
    const int l = (t & 31) * 2;
    __nv_bfloat16* vo = vcache + (((size_t)b * heads + hh) * 64 + d) * Lmax + pos0 + l0;
    if (l + 1 < nl && (((pos0 + l0) & 1) == 0)) {
      __nv_bfloat162 pr;
      pr.x = vt[d][l]; pr.y = vt[d][l + 1];
      *reinterpret_cast<__nv_bfloat162*>(vo + l) = pr;
    } else {
      if (l < nl) vo[l] = vt[d][l];
      if (l + 1 < nl) vo[l + 1] = vt[d][l + 1];
    }
  }
}

int rope_kv_launch(int dtype, const void* qkv, void* qout, void* kcache, void* vcache, void* vrows, int B, int Lq,
                   int heads, int Lmax, int pos0, const int* dpos, const float* cos_tab, const float* sin_tab, cudaStream_t st) {
  IVG_CHECK(pos0 + Lq <= Lmax, "rope_kv: pos0+Lq=%d exceeds cache length %d", pos0 + Lq, Lmax);
  if (B == 0 || Lq == 0) return 0;
  if (dtype == DT_BF16 && dpos == nullptr && Lq >= 64 && B <= 65535 && heads <= 65535) {      // prefill / teacher-forced pass
    rope_kv_tiled_kernel<<<dim3((Lq + 63) / 64, heads, B), 256, 0, st>>>(
        (const __nv_bfloat16*)qkv, (__nv_bfloat16*)qout, (__nv_bfloat16*)kcache, (__nv_bfloat16*)vcache, (__nv_bfloat16*)vrows, Lq,
        heads, Lmax, pos0, cos_tab, sin_tab);
    count_launch();
    IVG_LAUNCH_CHECK();
    return 0;
  }
  long long work = (long long)B * Lq * heads * 32;
  int blocks = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
  if (dtype == DT_BF16)
    rope_kv_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)qout,
                                                          (__nv_bfloat16*)kcache, (__nv_bfloat16*)vcache,
                                                          (__nv_bfloat16*)vrows, B, Lq, heads, Lmax, pos0, dpos,
                                                          cos_tab, sin_tab);
  else
    rope_kv_kernel<float><<<blocks, 256, 0, st>>>((const float*)qkv, (float*)qout, (float*)kcache, (float*)vcache,
                                                  (float*)vrows, B, Lq, heads, Lmax, pos0, dpos, cos_tab, sin_tab);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Row softmax of fp32 scores S[rows, lds] -> P[rows, ldp] (T).  causal: row r of every [Lq x Lk] block may
// see columns <= r + causal_off.  One warp per row; columns >= valid are written as zero up to ldp.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void softmax_kernel(const float* __restrict__ S, T* __restrict__ P, long long rows, int Lq, int Lk,
                               long long lds, long long ldp, int causal, int causal_off) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int rq = (int)(row % Lq);
  int valid = Lk;
  if (causal) { valid = rq + causal_off + 1; if (valid > Lk) valid = Lk; if (valid < 0) valid = 0; }
  const float* s = S + row * lds;
  float mx = -INFINITY;
  for (int c = lane; c < valid; c += 32) mx = fmaxf(mx, s[c]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float sum = 0.f;
  for (int c = lane; c < valid; c += 32) sum += __expf(s[c] - mx);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float inv = valid > 0 ? 1.0f / sum : 0.f;
  T* p = P + row * ldp;
  for (int c = lane; c < (int)ldp; c += 32) p[c] = from_f32<T>(c < valid ? __expf(s[c] - mx) * inv : 0.f);
}

int softmax_launch(int dtype, const float* S, void* P, long long rows, int Lq, int Lk, long long lds, long long ldp,
                   int causal, int causal_off, cudaStream_t st) {
  if (rows == 0) return 0;
  int blocks = (int)((rows + 7) / 8);
  if (dtype == DT_BF16)
    softmax_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(S, (__nv_bfloat16*)P, rows, Lq, Lk, lds, ldp, causal, causal_off);
  else
    softmax_kernel<float><<<blocks, 256, 0, st>>>(S, (float*)P, rows, Lq, Lk, lds, ldp, causal, causal_off);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Single-token decode attention over the KV cache (the per-token step of HF generate).  HBM-bound:
// streams K [Lcur,64] and V^T [64,Lcur] of one (batch, head) exactly once.  One CTA (4 warps) per (b, head):
//   phase 1: warps split positions, lanes hold 2 dims each -> scores in shared memory
//   phase 2: block softmax;  phase 3: warps split the 64 output dims, lanes stride positions (coalesced V^T rows)
// q: [B, heads, 1, 64] (already roped), out: [B, heads*64] T.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const T* __restrict__ q, const T* __restrict__ kcache, const T* __restrict__ vcache,
                   T* __restrict__ out, int heads, int Lmax, int Lcur, const int* __restrict__ dpos, float scale) {
  extern __shared__ float da_sm[];
  if (dpos) Lcur = *dpos + 1;  // [Lcur] scores, then 8 floats scratch
  float* sc = da_sm;
  float* red = da_sm + Lmax;
  const int bh = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* qp = q + (size_t)bh * 64;
  const float q0 = to_f32(qp[2 * lane]) * scale, q1 = to_f32(qp[2 * lane + 1]) * scale;
  const T* kp = kcache + (size_t)bh * Lmax * 64;
  for (int l = warp; l < Lcur; l += 4) {
    const T* kr = kp + (size_t)l * 64;
    float s = q0 * to_f32(kr[2 * lane]) + q1 * to_f32(kr[2 * lane + 1]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) sc[l] = s;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int l = threadIdx.x; l < Lcur; l += 128) mx = fmaxf(mx, sc[l]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float sum = 0.f;
  for (int l = threadIdx.x; l < Lcur; l += 128) { float e = __expf(sc[l] - mx); sc[l] = e; sum += e; }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) red[4 + warp] = sum;
  __syncthreads();
  const float inv = 1.0f / (red[4] + red[5] + red[6] + red[7]);
  const T* vp = vcache + (size_t)bh * 64 * Lmax;
  const int b = bh / heads, hh = bh - b * heads;
  for (int d = warp; d < 64; d += 4) {
    const T* vr = vp + (size_t)d * Lmax;
    float a = 0.f;
    for (int l = lane; l < Lcur; l += 32) a = fmaf(sc[l], to_f32(vr[l]), a);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) out[((size_t)b * heads + hh) * 64 + d] = from_f32<T>(a * inv);
  }
}

int decode_attn_launch(int dtype, const void* q, const void* kcache, const void* vcache, void* out, int B, int heads,
                       int Lmax, int Lcur, const int* dpos, float scale, cudaStream_t st) {
  IVG_CHECK(dpos || (Lcur >= 1 && Lcur <= Lmax), "decode_attn: Lcur=%d out of range (Lmax=%d)", Lcur, Lmax);
  if (B == 0) return 0;
  size_t smem = (size_t)(Lmax + 8) * sizeof(float);
  if (dtype == DT_BF16)
    decode_attn_kernel<__nv_bfloat16><<<B * heads, 128, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)kcache,
                                                                    (const __nv_bfloat16*)vcache, (__nv_bfloat16*)out,
                                                                    heads, Lmax, Lcur, dpos, scale);
  else
    decode_attn_kernel<float><<<B * heads, 128, smem, st>>>((const float*)q, (const float*)kcache, (const float*)vcache,
                                                            (float*)out, heads, Lmax, Lcur, dpos, scale);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused decode-step attention: RoPE(q,k) + KV-cache append + attention over the cache, one CTA per (batch, head).
// This is the per-token rollout step of HF generate (predict.py:64) with an incremental KV cache.  HBM-bound:
// K rows [Lcur,64] and V^T rows [64,Lcur] of the (b,h) slab are streamed exactly once with 16-byte coalesced
// loads (8 lanes per K row; a full warp per V^T row), everything else lives in shared memory.
//   qkv [B, 3*hidden] (T) for the current token, position pos = *dpos (or pos_arg), Lcur = pos + 1.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[16 / sizeof(T)]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  if constexpr (sizeof(T) == 2) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 v = __bfloat1622float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
  } else {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_attn_fused_kernel(const T* __restrict__ qkv, T* __restrict__ kcache, T* __restrict__ vcache,
                         T* __restrict__ out, int heads, int Lmax, int pos_arg, const int* __restrict__ dpos,
                         const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, float scale) {
  constexpr int VN = Vec16<T>::N;            // elements per 16-byte load (8 bf16 / 4 fp32)
  constexpr int LPR = 64 / VN;               // lanes per K row (8 / 16)
  extern __shared__ float fa_sm[];           // [Lmax + VN] probabilities, then q[64], red[16]
  pdl_wait();
  pdl_launch_dependents();
  const int pos = dpos ? *dpos : pos_arg;
  const int Lcur = pos + 1;
  float* sc = fa_sm;
  float* qs = fa_sm + Lmax + VN;
  float* red = qs + 64;
  const int bh = blockIdx.x;
  const int b = bh / heads, hh = bh - b * heads;
  const int Hd = heads * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* kslab = kcache + (size_t)bh * Lmax * 64;
  T* vslab = vcache + (size_t)bh * 64 * Lmax;

  if (tid < 32) {                            // RoPE + append (one warp: pair (i, i+32))
    const T* row = qkv + (size_t)b * 3 * Hd;
    const float cs = __ldg(cos_tab + (size_t)pos * 32 + tid), sn = __ldg(sin_tab + (size_t)pos * 32 + tid);
    const float q0 = to_f32(row[hh * 64 + tid]), q1 = to_f32(row[hh * 64 + tid + 32]);
    const float k0 = to_f32(row[Hd + hh * 64 + tid]), k1 = to_f32(row[Hd + hh * 64 + tid + 32]);
    // q is rounded to T exactly like the unfused path (rope_kv_kernel) so both decode paths agree bit for bit
    qs[tid] = to_f32(from_f32<T>(q0 * cs - q1 * sn)) * scale;
    qs[tid + 32] = to_f32(from_f32<T>(q1 * cs + q0 * sn)) * scale;
    kslab[(size_t)pos * 64 + tid] = from_f32<T>(k0 * cs - k1 * sn);
    kslab[(size_t)pos * 64 + tid + 32] = from_f32<T>(k1 * cs + k0 * sn);
    vslab[(size_t)tid * Lmax + pos] = row[2 * Hd + hh * 64 + tid];
    vslab[(size_t)(tid + 32) * Lmax + pos] = row[2 * Hd + hh * 64 + tid + 32];
  }
  __syncthreads();                           // orders the global K/V writes above for this CTA, publishes qs

  // ---- scores: LPR lanes per cache row ----
  const int sub = tid % LPR, rslot = tid / LPR;
  constexpr int ROWS_PER_PASS = 256 / LPR;
  float qreg[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) qreg[i] = qs[sub * VN + i];
  for (int l0 = 0; l0 < Lcur; l0 += ROWS_PER_PASS) {
    const int l = l0 + rslot;
    float part = 0.f;
    if (l < Lcur) {
      float kv[VN];
      load8<T>(kslab + (size_t)l * 64 + sub * VN, kv);
#pragma unroll
      for (int i = 0; i < VN; ++i) part = fmaf(qreg[i], kv[i], part);
    }
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (sub == 0 && l < Lcur) sc[l] = part;
  }
  __syncthreads();
  // ---- softmax over sc[0..Lcur) ----
  float mx = -INFINITY;
  for (int l = tid; l < Lcur; l += 256) mx = fmaxf(mx, sc[l]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  float sum = 0.f;
  for (int l = tid; l < Lcur + VN; l += 256) {
    const float e = l < Lcur ? __expf(sc[l] - mx) : 0.f;
    sc[l] = e;                               // zero tail so the vector loop below can run past Lcur
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
  // ---- out[d] = sum_l p[l] * V^T[d][l] : warp per row, 16-byte loads along l ----
  for (int d = warp; d < 64; d += 8) {
    const T* vr = vslab + (size_t)d * Lmax;
    float a = 0.f;
    for (int l = lane * VN; l < Lcur; l += 32 * VN) {
      float vv[VN];
      load8<T>(vr + l, vv);
#pragma unroll
      for (int i = 0; i < VN; ++i) a += (l + i < Lcur) ? sc[l + i] * vv[i] : 0.f;   // cache past Lcur is uninitialised
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) out[(size_t)b * Hd + hh * 64 + d] = from_f32<T>(a * inv);
  }
}

int decode_attn_fused_launch(int dtype, const void* qkv, void* kcache, void* vcache, void* out, int B, int heads,
                             int Lmax, int pos, const int* dpos, const float* cos_tab, const float* sin_tab,
                             float scale, cudaStream_t st) {
  IVG_CHECK(dpos || (pos >= 0 && pos < Lmax), "decode_attn_fused: pos=%d out of range (Lmax=%d)", pos, Lmax);
  IVG_CHECK(Lmax % 8 == 0, "decode_attn_fused: Lmax must be a multiple of 8");
  if (B == 0) return 0;
  size_t smem = (size_t)(Lmax + 8 + 64 + 16) * sizeof(float);
  if (dtype == DT_BF16)
    IVG_CUDA(launch_k(decode_attn_fused_kernel<__nv_bfloat16>, dim3(B * heads), dim3(256), smem, st,
                      (const __nv_bfloat16*)qkv, (__nv_bfloat16*)kcache, (__nv_bfloat16*)vcache, (__nv_bfloat16*)out,
                      heads, Lmax, pos, dpos, cos_tab, sin_tab, scale));
  else
    IVG_CUDA(launch_k(decode_attn_fused_kernel<float>, dim3(B * heads), dim3(256), smem, st, (const float*)qkv,
                      (float*)kcache, (float*)vcache, (float*)out, heads, Lmax, pos, dpos, cos_tab, sin_tab, scale));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Greedy argmax over fp32 logits rows (first maximal index, torch.argmax rule).  One CTA per row.
// ---------------------------------------------------------------------------------------------
__global__ void argmax_kernel(const float* __restrict__ logits, long long ld, int V, long long* __restrict__ out,
                              long long out_stride, const int* __restrict__ dpos) {
  __shared__ float sv[32];
  __shared__ int si[32];
  pdl_wait();
  pdl_launch_dependents();
  if (dpos) out += *dpos + 1;
  const float* row = logits + (size_t)blockIdx.x * ld;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float v = row[c];
    if (v > bv || (v == bv && c < bi)) { bv = v; bi = c; }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, off);
    int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    bv = lane < nw ? sv[lane] : -INFINITY; bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) out[(size_t)blockIdx.x * out_stride] = bi;
  }
}

int argmax_launch(const float* logits, long long ld, int rows, int V, long long* out, long long out_stride,
                  const int* dpos, cudaStream_t st) {
  if (rows == 0) return 0;
  IVG_CUDA(launch_k(argmax_kernel, dim3(rows), dim3(256), 0, st, logits, ld, V, out, out_stride, dpos));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Top-k sampling (HF TopKLogitsWarper + softmax + multinomial; predict.py:58-63 uses top_k=100, T=1):
// keep logits >= the k-th largest (ties keep extras), softmax(logits / temperature), draw one sample by
// inverse CDF in index order from a counter-based uniform (seed, step, row).  One CTA per row; the k-th
// largest value is found by a 4-pass radix select on order-preserving keys held in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fkey(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long step, unsigned long long row) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (step * 0x100000001B3ull + row + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)((z >> 40) + 0.5) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(512)
topk_sample_kernel(const float* __restrict__ logits, long long ld, int V, int k, float inv_temp,
                   unsigned long long seed, unsigned long long step, long long* __restrict__ out,
                   long long out_stride, const int* __restrict__ dpos,
                   const unsigned long long* __restrict__ dseed) {
  extern __shared__ uint32_t tk_sm[];
  pdl_wait();
  pdl_launch_dependents();
  if (dpos) { out += *dpos + 1; step += (unsigned long long)*dpos; }
  if (dseed) seed += *dseed;  // [V] keys, then 256 hist, then scratch
  uint32_t* keys = tk_sm;
  uint32_t* hist = tk_sm + V;
  __shared__ uint32_t s_prefix, s_remaining;
  __shared__ float s_red[32];
  const float* row = logits + (size_t)blockIdx.x * ld;
  for (int c = threadIdx.x; c < V; c += blockDim.x) keys[c] = fkey(row[c] * inv_temp);
  if (threadIdx.x == 0) { s_prefix = 0; s_remaining = (uint32_t)(k < V ? k : V); }
  __syncthreads();
  // radix select: find key of the k-th largest
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
      uint32_t kk = keys[c];
      if ((kk & mask) == prefix) atomicAdd(&hist[(kk >> shift) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t rem = s_remaining;
      int b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      s_prefix = prefix | ((uint32_t)b << shift);
      s_remaining = rem;
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;
  // max is the largest key -> recover float max for a stable softmax
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += blockDim.x) if (keys[c] >= kth) mx = fmaxf(mx, row[c] * inv_temp);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < nw; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  // Inverse-CDF draw over the survivors, in index order, parallel and deterministic: thread t owns the contiguous
  // index chunk [t*cpt, (t+1)*cpt); block-wide inclusive scan of the chunk masses; the FIRST thread whose cumulative
  // mass reaches u (block-wide min) walks its <= cpt elements.  Values are recovered from the shared-memory keys.
  {
    __shared__ float s_hi[512];
    __shared__ int s_win, s_lastmass;
    if (threadIdx.x == 0) { s_win = 0x7fffffff; s_lastmass = 0; }
    const int cpt = (V + blockDim.x - 1) / blockDim.x;
    const int c0 = threadIdx.x * cpt, c1 = min(V, c0 + cpt);
    float local = 0.f;
    for (int c = c0; c < c1; ++c) {
      const uint32_t kk = keys[c];
      if (kk >= kth) {
        const float f = __uint_as_float((kk & 0x80000000u) ? (kk ^ 0x80000000u) : ~kk);
        local += __expf(f - mx);
      }
    }
    float incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) s_red[warp] = incl;
    __syncthreads();
    float wbase = 0.f, total = 0.f;
    for (int i = 0; i < nw; ++i) { const float v = s_red[i]; if (i < warp) wbase += v; total += v; }
    const float hi = wbase + incl;
    s_hi[threadIdx.x] = hi;
    const float u = hash_uniform(seed, step, blockIdx.x) * total;
    if (local > 0.f) {
      atomicMax(&s_lastmass, (int)threadIdx.x);
      if (hi >= u) atomicMin(&s_win, (int)threadIdx.x);
    }
    __syncthreads();
    const int win = (s_win == 0x7fffffff) ? s_lastmass : s_win;
    if ((int)threadIdx.x == win) {
      float acc = win > 0 ? s_hi[win - 1] : 0.f;
      int pick = -1, last = c0;
      for (int c = c0; c < c1; ++c) {
        const uint32_t kk = keys[c];
        if (kk >= kth) {
          const float f = __uint_as_float((kk & 0x80000000u) ? (kk ^ 0x80000000u) : ~kk);
          acc += __expf(f - mx);
          last = c;
          if (acc >= u) { pick = c; break; }
        }
      }
      out[(size_t)blockIdx.x * out_stride] = pick >= 0 ? pick : last;
    }
  }
}

int topk_sample_launch(const float* logits, long long ld, int rows, int V, int k, float temperature,
                       unsigned long long seed, unsigned long long step, long long* out, long long out_stride,
                       const int* dpos, const unsigned long long* dseed, cudaStream_t st) {
  IVG_CHECK(k >= 1 && temperature > 0.f, "topk_sample: bad k=%d / temperature=%f", k, temperature);
  if (rows == 0) return 0;
  size_t smem = (size_t)(V + 256) * sizeof(uint32_t);
  static PerDeviceOnce attr_once;
  if (attr_once.pending()) {
    IVG_CUDA(cudaFuncSetAttribute(topk_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.mark();
  }
  IVG_CHECK(smem <= 200 * 1024, "topk_sample: vocab %d too large for the shared-memory select", V);
  IVG_CUDA(launch_k(topk_sample_kernel, dim3(rows), dim3(512), smem, st, logits, ld, V, k, 1.0f / temperature, seed, step,
                    out, out_stride, dpos, dseed));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Shifted cross-entropy (LlamaForCausalLM loss: logits[:, :-1] vs labels[:, 1:], ignore_index -100, mean over
// labelled positions; train_gpt.py:792).  logits [B, L, ld] fp32, labels [B, L] int64.  One CTA per (b, l<L-1):
// loss_rows[b*(L-1)+l] = logsumexp(logits[b,l]) - logits[b,l,labels[b,l+1]], or 0 with valid=0 when the label
// is ignore_index.  A second 1-CTA kernel reduces sum / count deterministically (fp64).
// ---------------------------------------------------------------------------------------------
__global__ void ce_rows_kernel(const float* __restrict__ logits, long long ld, int V, int L,
                               const long long* __restrict__ labels, float* __restrict__ loss_rows,
                               float* __restrict__ valid) {
  __shared__ float s_red[32];
  const int b = blockIdx.x / (L - 1), l = blockIdx.x % (L - 1);
  const long long t = labels[(size_t)b * L + l + 1];
  if (t < 0 || t >= V) {  // ignore_index (-100) or out-of-vocabulary label
    if (threadIdx.x == 0) { loss_rows[blockIdx.x] = 0.f; valid[blockIdx.x] = 0.f; }
    return;
  }
  const float* row = logits + ((size_t)b * L + l) * ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, row[c]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < nw; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) sum += expf(row[c] - mx);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < nw; ++i) tot += s_red[i];
    loss_rows[blockIdx.x] = (logf(tot) + mx) - row[t];
    valid[blockIdx.x] = 1.f;
  }
}

__global__ void ce_mean_kernel(const float* __restrict__ v, const float* __restrict__ valid, int n,
                               float* __restrict__ out) {
  __shared__ double s[256], c[256];
  double a = 0.0, k = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { a += (double)v[i]; k += (double)valid[i]; }
  s[threadIdx.x] = a; c[threadIdx.x] = k;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) { s[threadIdx.x] += s[threadIdx.x + o]; c[threadIdx.x] += c[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = c[0] > 0 ? (float)(s[0] / c[0]) : 0.f; out[1] = (float)c[0]; }
}

int ce_loss_launch(const float* logits, long long ld, int B, int L, int V, const long long* labels, float* loss_rows,
                   float* valid_ws, float* loss_out /* [2] = mean, count */, cudaStream_t st) {
  IVG_CHECK(L >= 2, "ce_loss: sequence length must be >= 2");
  if (B == 0) return 0;
  const int rows = B * (L - 1);
  ce_rows_kernel<<<rows, 256, 0, st>>>(logits, ld, V, L, labels, loss_rows, valid_ws);
  ce_mean_kernel<<<1, 256, 0, st>>>(loss_rows, valid_ws, rows, loss_out);
  count_launch(2);
  IVG_LAUNCH_CHECK();
  return 0;
}


// x[m,:] += extra[m,:]   (fp32 residual stream; used for action embeddings)
__global__ void add_rows_kernel(float* __restrict__ x, const float* __restrict__ e, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] += e[i];
}
int add_rows_launch(float* x, const float* e, long long n, cudaStream_t st) {
  if (n == 0) return 0;
  int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  add_rows_kernel<<<blocks, 256, 0, st>>>(x, e, n);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

__global__ void incr_kernel(int* p, int by) {
  pdl_wait();
  pdl_launch_dependents();
  *p += by;
}
int incr_launch(int* p, int by, cudaStream_t st) {
  IVG_CUDA(launch_k(incr_kernel, dim3(1), dim3(1), 0, st, p, by));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Forced "slot" positions of an action-conditioned rollout (reference ivideogpt/transformer/action_model.py:78-114):
// every `period`-th position from `slot0` holds the separator token, never a sampled one, and the embedding fed at
// slot i gets action_linear(a_i) added (:80-81).  Both kernels read the position from device memory (graph-safe).
//   slot_embed_add : x[b, :] += slot_emb[b, i, :]   if the position being fed, *dpos, is slot i
//   slot_force     : tokens[b, *dpos + 1] = token   if position *dpos + 1 is a slot
// ---------------------------------------------------------------------------------------------
__global__ void slot_embed_add_kernel(float* __restrict__ x, const float* __restrict__ slot_emb, const int* __restrict__ dpos,
                                      int B, int Hd, int slot0, int period, int nslots) {
  pdl_wait();
  pdl_launch_dependents();
  const int pos = *dpos;
  if (pos < slot0 || (pos - slot0) % period != 0) return;
  const int i = (pos - slot0) / period;
  if (i >= nslots) return;
  const int hv = Hd / 4;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < B * hv; t += gridDim.x * blockDim.x) {
    const int b = t / hv, c = t - b * hv;
    float4 v = reinterpret_cast<float4*>(x)[t];
    const float4 e = __ldg(reinterpret_cast<const float4*>(slot_emb + ((size_t)b * nslots + i) * Hd) + c);
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    reinterpret_cast<float4*>(x)[t] = v;
  }
}
int slot_embed_add_launch(float* x, const float* slot_emb, const int* dpos, int B, int Hd, int slot0, int period,
                          int nslots, cudaStream_t st) {
  IVG_CHECK(Hd % 4 == 0 && period >= 1 && nslots >= 1 && dpos != nullptr, "slot_embed_add: bad arguments");
  const int work = B * (Hd / 4);
  IVG_CUDA(launch_k(slot_embed_add_kernel, dim3((work + 255) / 256), dim3(256), 0, st, x, slot_emb, dpos, B, Hd, slot0,
                    period, nslots));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

__global__ void slot_force_kernel(long long* __restrict__ tokens, long long stride, const int* __restrict__ dpos, int B,
                                  int slot0, int period, long long token) {
  pdl_wait();
  pdl_launch_dependents();
  const int nxt = *dpos + 1;
  if (nxt < slot0 || (nxt - slot0) % period != 0) return;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) tokens[(size_t)b * stride + nxt] = token;
}
int slot_force_launch(long long* tokens, long long stride, const int* dpos, int B, int slot0, int period, long long token,
                      cudaStream_t st) {
  IVG_CHECK(period >= 1 && dpos != nullptr, "slot_force: bad arguments");
  IVG_CUDA(launch_k(slot_force_kernel, dim3((B + 127) / 128), dim3(128), 0, st, tokens, stride, dpos, B, slot0, period, token));
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

}  // namespace ivg
