// Backward-pass kernels of the Llama-style transformer (train_gpt.py:792-804: model(input_ids, labels).loss ->
// accelerator.backward -> AdamW).  All contractions (dgrad / wgrad / attention backward) run on the tcgen05 GEMM of
// gemm_tc.cu; this file holds what is left: batched transposes (the GEMM wants both operands K-major), SwiGLU / RMSNorm /
// softmax / RoPE backward, cross-entropy gradient, embedding scatter-add and the fused AdamW update.
// Residual-stream gradients are fp32; GEMM operands are T (bf16 or fp32->TF32).
#include "common.cuh"

namespace ivg {

// ---------------------------------------------------------------------------------------------
// batched 2-D transpose: in [batch][R][C] (row pitch ld_in) -> out [batch][C][R] (row pitch ld_out)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int Cc, long long ld_in,
                                 long long ld_out, long long bs_in, long long bs_out) {
  __shared__ T tile[32][33];
  const T* src = in + (size_t)blockIdx.z * bs_in;
  T* dst = out + (size_t)blockIdx.z * bs_out;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = src[(size_t)r * ld_in + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) dst[(size_t)c * ld_out + r] = tile[threadIdx.x][i];
  }
}

int transpose_launch(int dtype, const void* in, void* out, int batch, int R, int Cc, long long ld_in, long long ld_out,
                     long long bs_in, long long bs_out, cudaStream_t st) {
  if (batch <= 0 || R <= 0 || Cc <= 0) return 0;
  IVG_CHECK(batch <= 65535 && (R + 31) / 32 <= 65535, "transpose: grid too large (batch %d, rows %d)", batch, R);
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, batch), block(32, 8);
  if (dtype == DT_BF16)
    transpose_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, Cc, ld_in,
                                                            ld_out, bs_in, bs_out);
  else
    transpose_kernel<float><<<grid, block, 0, st>>>((const float*)in, (float*)out, R, Cc, ld_in, ld_out, bs_in, bs_out);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// SwiGLU on an interleaved gate/up buffer gu [M, 2I] (col 2j = gate_j, 2j+1 = up_j)
//   fwd: act[m,j] = silu(g) * u          bwd: d_g = d_act * u * s(g) * (1 + g (1 - s(g))),  d_u = d_act * silu(g)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void swiglu_fwd_kernel(const T* __restrict__ gu, T* __restrict__ act, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = to_f32(gu[2 * i]), u = to_f32(gu[2 * i + 1]);
    act[i] = from_f32<T>(silu_f(g) * u);
  }
}
template <typename T>
__global__ void swiglu_bwd_kernel(const T* __restrict__ gu, const T* __restrict__ dact, T* __restrict__ dgu,
                                  long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = to_f32(gu[2 * i]), u = to_f32(gu[2 * i + 1]), d = to_f32(dact[i]);
    const float s = 1.0f / (1.0f + __expf(-g));
    dgu[2 * i] = from_f32<T>(d * u * s * (1.0f + g * (1.0f - s)));
    dgu[2 * i + 1] = from_f32<T>(d * g * s);
  }
}

int swiglu_launch(int dtype, int backward, const void* gu, const void* dact, void* out, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (dtype == DT_BF16) {
    if (backward) swiglu_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)gu, (const __nv_bfloat16*)dact, (__nv_bfloat16*)out, n);
    else swiglu_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)gu, (__nv_bfloat16*)out, n);
  } else {
    if (backward) swiglu_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)gu, (const float*)dact, (float*)out, n);
    else swiglu_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)gu, (float*)out, n);
  }
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// RMSNorm backward.  y = w * x * r, r = rsqrt(mean(x^2) + eps).  One warp per row:
//   dres[m,:] += r * (w * dy) - x * r^3 / H * sum_j(w_j dy_j x_j)            (fp32, accumulated in place)
//   dw partial: each CTA (8 rows) writes dw_part[block][:] = sum_rows dy * x * r; a second kernel reduces blocks.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rmsnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const T* __restrict__ dy,
                                   float* __restrict__ dres, float* __restrict__ dw_part, long long M, int H, float eps) {
  extern __shared__ float rb_sm[];   // [H] CTA-level dw accumulator
  for (int i = threadIdx.x; i < H; i += blockDim.x) rb_sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long row = (long long)blockIdx.x * nw + warp;
  if (row < M) {
    const float* xr = x + row * H;
    const T* dr = dy + row * H;
    float ss = 0.f, dot = 0.f;
    for (int i = lane; i < H; i += 32) {
      const float xv = xr[i];
      ss = fmaf(xv, xv, ss);
      dot = fmaf(w[i] * to_f32(dr[i]), xv, dot);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
      dot += __shfl_xor_sync(0xffffffffu, dot, off);
    }
    const float r = rsqrtf(ss / (float)H + eps);
    const float coef = dot * r * r * r / (float)H;
    float* gr = dres + row * H;
    for (int i = lane; i < H; i += 32) {
      const float xv = xr[i], d = to_f32(dr[i]);
      gr[i] += r * w[i] * d - xv * coef;
      atomicAdd(&rb_sm[i], d * xv * r);   // 8 warps of one CTA: shared-memory atomics, order-dependent in the last ulp
    }
  }
  __syncthreads();
  float* dst = dw_part + (size_t)blockIdx.x * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) dst[i] = rb_sm[i];
}

__global__ void colsum_kernel(const float* __restrict__ part, float* __restrict__ out, int nblocks, int H, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += part[(size_t)b * H + i];
  out[i] = accumulate ? out[i] + s : s;
}

int rmsnorm_bwd_launch(int dtype, const float* x, const float* w, const void* dy, float* dres, float* dw_part, float* dw,
                       long long M, int H, float eps, cudaStream_t st) {
  if (M <= 0) return 0;
  const int nblocks = (int)((M + 7) / 8);
  if (dtype == DT_BF16)
    rmsnorm_bwd_kernel<__nv_bfloat16><<<nblocks, 256, H * sizeof(float), st>>>(x, w, (const __nv_bfloat16*)dy, dres, dw_part, M, H, eps);
  else
    rmsnorm_bwd_kernel<float><<<nblocks, 256, H * sizeof(float), st>>>(x, w, (const float*)dy, dres, dw_part, M, H, eps);
  colsum_kernel<<<(H + 255) / 256, 256, 0, st>>>(dw_part, dw, nblocks, H, 0);
  count_launch(2);
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// softmax backward (causal): dS = scale * P * (dP - sum_j P_j dP_j); columns beyond the causal limit are zero.
// P [rows, ld] (T), dP [rows, ld] fp32, dS [rows, ld] (T).  One warp per row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void softmax_bwd_kernel(const T* __restrict__ P, const float* __restrict__ dP, T* __restrict__ dS,
                                   long long rows, int Lq, int Lk, long long ld, int causal, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int rq = (int)(row % Lq);
  int valid = Lk;
  if (causal) { valid = rq + 1; if (valid > Lk) valid = Lk; }
  const T* p = P + row * ld;
  const float* dp = dP + row * ld;
  float s = 0.f;
  for (int c = lane; c < valid; c += 32) s = fmaf(to_f32(p[c]), dp[c], s);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  T* o = dS + row * ld;
  for (int c = lane; c < (int)ld; c += 32) o[c] = from_f32<T>(c < valid ? scale * to_f32(p[c]) * (dp[c] - s) : 0.f);
}

int softmax_bwd_launch(int dtype, const void* P, const float* dP, void* dS, long long rows, int Lq, int Lk, long long ld,
                       int causal, float scale, cudaStream_t st) {
  if (rows <= 0) return 0;
  int blocks = (int)((rows + 7) / 8);
  if (dtype == DT_BF16)
    softmax_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)P, dP, (__nv_bfloat16*)dS, rows, Lq, Lk, ld, causal, scale);
  else
    softmax_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)P, dP, (float*)dS, rows, Lq, Lk, ld, causal, scale);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// RoPE backward + re-assembly of the fused qkv gradient:
//   dq', dk' [B, heads, L, 64] (fp32, gradients w.r.t. the ROTATED q / k), dv [B, heads, L, 64] (fp32)
//   -> dqkv [B*L, 3*hidden] (T):  dq1 = dq'1 cos + dq'2 sin, dq2 = -dq'1 sin + dq'2 cos  (same for k), dv copied.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rope_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
                                T* __restrict__ dqkv, int B, int L, int heads, const float* __restrict__ cos_tab,
                                const float* __restrict__ sin_tab) {
  const long long total = (long long)B * L * heads * 32;
  const int Hd = heads * 64;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx & 31);
    long long r = idx >> 5;
    const int hh = (int)(r % heads); r /= heads;
    const int l = (int)(r % L);
    const int b = (int)(r / L);
    const float cs = __ldg(cos_tab + (size_t)l * 32 + i), sn = __ldg(sin_tab + (size_t)l * 32 + i);
    const size_t src = (((size_t)b * heads + hh) * L + l) * 64;
    T* row = dqkv + ((size_t)b * L + l) * (3 * Hd);
    const float q1 = dq[src + i], q2 = dq[src + i + 32];
    row[hh * 64 + i] = from_f32<T>(q1 * cs + q2 * sn);
    row[hh * 64 + i + 32] = from_f32<T>(-q1 * sn + q2 * cs);
    const float k1 = dk[src + i], k2 = dk[src + i + 32];
    row[Hd + hh * 64 + i] = from_f32<T>(k1 * cs + k2 * sn);
    row[Hd + hh * 64 + i + 32] = from_f32<T>(-k1 * sn + k2 * cs);
    row[2 * Hd + hh * 64 + i] = from_f32<T>(dv[src + i]);
    row[2 * Hd + hh * 64 + i + 32] = from_f32<T>(dv[src + i + 32]);
  }
}

int rope_bwd_launch(int dtype, const float* dq, const float* dk, const float* dv, void* dqkv, int B, int L, int heads,
                    const float* cos_tab, const float* sin_tab, cudaStream_t st) {
  if (B <= 0 || L <= 0) return 0;
  long long work = (long long)B * L * heads * 32;
  int blocks = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  if (dtype == DT_BF16)
    rope_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(dq, dk, dv, (__nv_bfloat16*)dqkv, B, L, heads, cos_tab, sin_tab);
  else
    rope_bwd_kernel<float><<<blocks, 256, 0, st>>>(dq, dk, dv, (float*)dqkv, B, L, heads, cos_tab, sin_tab);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Cross-entropy gradient (shifted labels, mean over labelled positions): for row (b, l < L-1) with label
// t = labels[b, l+1] >= 0:  dlogits[b,l,:] = (softmax(logits[b,l,:]) - onehot(t)) * gscale / count; other rows zero.
// count is read from device memory (loss_out[1] of ivgpt_ce_loss).  One CTA per row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void ce_bwd_kernel(const float* __restrict__ logits, long long ld, int V, int L,
                              const long long* __restrict__ labels, const float* __restrict__ count, float gscale,
                              T* __restrict__ dlogits, long long ldd) {
  __shared__ float s_red[32];
  const int b = blockIdx.x / L, l = blockIdx.x % L;
  T* out = dlogits + ((size_t)b * L + l) * ldd;
  long long t = -1;
  if (l < L - 1) t = labels[(size_t)b * L + l + 1];
  if (t < 0 || t >= V) {
    for (int c = threadIdx.x; c < (int)ldd; c += blockDim.x) out[c] = from_f32<T>(0.f);
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
  float tot = 0.f;
  for (int i = 0; i < nw; ++i) tot += s_red[i];
  const float k = gscale / fmaxf(count[0], 1.0f);
  const float inv = 1.0f / tot;
  for (int c = threadIdx.x; c < (int)ldd; c += blockDim.x) {
    float g = 0.f;
    if (c < V) g = (expf(row[c] - mx) * inv - (c == (int)t ? 1.0f : 0.0f)) * k;
    out[c] = from_f32<T>(g);
  }
}

int ce_bwd_launch(int dtype, const float* logits, long long ld, int B, int L, int V, const long long* labels,
                  const float* count, float gscale, void* dlogits, long long ldd, cudaStream_t st) {
  if (B <= 0) return 0;
  if (dtype == DT_BF16)
    ce_bwd_kernel<__nv_bfloat16><<<B * L, 256, 0, st>>>(logits, ld, V, L, labels, count, gscale, (__nv_bfloat16*)dlogits, ldd);
  else
    ce_bwd_kernel<float><<<B * L, 256, 0, st>>>(logits, ld, V, L, labels, count, gscale, (float*)dlogits, ldd);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Embedding backward: dE[ids[m], :] += dx[m, :]   (fp32 atomics; dE must be zero-initialised by the caller)
// ---------------------------------------------------------------------------------------------
__global__ void embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dx, float* __restrict__ dE,
                                 long long M, int H, long long vocab) {
  const long long total = M * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / H;
    const int c = (int)(i - m * H);
    long long id = ids[m];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    atomicAdd(dE + id * H + c, dx[i]);
  }
}

int embed_bwd_launch(const long long* ids, const float* dx, float* dE, long long M, int H, long long vocab,
                     cudaStream_t st) {
  if (M <= 0) return 0;
  long long work = M * H;
  int blocks = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  embed_bwd_kernel<<<blocks, 256, 0, st>>>(ids, dx, dE, M, H, vocab);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused AdamW (torch.optim.AdamW semantics, train_gpt.py:648-658): decoupled weight decay, bias correction.
//   p -= lr * wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// g is multiplied by gscale first (gradient averaging over ranks / clipping factor).
// ---------------------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, float gscale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float pi = p[i];
    pi -= lr * wd * pi;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
  }
}

int adamw_launch(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                 float wd, int step, float gscale, cudaStream_t st) {
  if (n <= 0) return 0;
  const float bc1 = 1.0f - powf(b1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(b2, (float)step));
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  adamw_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// y (fp32) += x (T)  : gradient accumulation into the fp32 residual-stream gradient
template <typename T>
__global__ void add_to_f32_kernel(float* __restrict__ y, const T* __restrict__ x, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += to_f32(x[i]);
}
int add_to_f32_launch(int dtype, float* y, const void* x, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (dtype == DT_BF16) add_to_f32_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(y, (const __nv_bfloat16*)x, n);
  else add_to_f32_kernel<float><<<blocks, 256, 0, st>>>(y, (const float*)x, n);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Attention dropout (HF LlamaAttention: nn.functional.dropout(attn_weights, p=attention_dropout, training=True);
// reference scripts/pretrain/oxe-64-act-free.sh:31 trains with 0.1).  y[i] = keep(i) ? x[i] / (1 - p) : 0 with a
// counter-based generator: keep(i) depends only on (seed, element index), so the backward pass regenerates the mask of
// the forward pass from the seed instead of storing it (dP is masked in place, P' = dropout(P) is recomputed for dV).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dropout_uniform(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);      // [0, 1)
}
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float p, float inv_keep,
                               unsigned long long seed) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = dropout_uniform(seed, (unsigned long long)i) >= p ? from_f32<T>(to_f32(x[i]) * inv_keep) : from_f32<T>(0.f);
}
int dropout_launch(int dtype, const void* x, void* y, long long n, float p, unsigned long long seed, cudaStream_t st) {
  IVG_CHECK(p >= 0.f && p < 1.f, "dropout: p=%f must be in [0, 1)", p);
  if (n <= 0) return 0;
  const float inv_keep = 1.0f / (1.0f - p);
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (dtype == DT_BF16) dropout_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, p, inv_keep, seed);
  else dropout_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, n, p, inv_keep, seed);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

}  // namespace ivg
