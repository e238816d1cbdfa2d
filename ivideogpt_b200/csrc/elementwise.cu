// Memory-bound kernels of the ctx_vqgan tokenizer path: GroupNorm statistics / apply (+SiLU, +pos-emb),
// first/last 3-channel convolutions on CUDA cores, nearest-2x upsample, (de)patchify, codebook gathers and
// token (de)serialisation.  All activations are NHWC ([N, H, W, C], C fastest) in fp32 or bf16.
// Reference call sites are cited per kernel.
#include "common.cuh"

namespace ivg {

// ---------------------------------------------------------------------------------------------
// GroupNorm (torch.nn.GroupNorm semantics; eps 1e-6 in ResnetBlock2D/mid/out norms, vae.py:110,133,
// eps 1e-5 in CrossAttentionBlock q/kv norms, conditional_vae.py:26-27).
// Two deterministic stages: per-(sample, slab) partial sums in fp32 -> per-(sample, group) mean/rstd
// combined in fp64.  "sample" may span several frames (kv_norm normalises both context frames jointly,
// conditional_vae.py:41-44) -- the caller just passes rows = frames*H*W.
// ---------------------------------------------------------------------------------------------
constexpr int GN_SLAB_ROWS = 64;  // pixels per partial-sum CTA

// 16-byte vector of T -> 8 floats (bf16) or 4 floats (fp32)
template <typename T> struct GnVec { static constexpr int N = 16 / sizeof(T); };
template <typename T>
__device__ __forceinline__ void gn_load(const T* p, float (&f)[16 / sizeof(T)]) {
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
__device__ __forceinline__ void gn_store(T* p, const float (&f)[16 / sizeof(T)]) {
  uint4 u;
  if constexpr (sizeof(T) == 2) {
    u = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  } else {
    u = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  *reinterpret_cast<uint4*>(p) = u;
}

constexpr int GN_PART_ROWS = 256;  // pixels per CTA of the stand-alone statistics pass (64 KB of bf16 at C = 128)

template <typename T>
__global__ void gn_partial_kernel(const T* __restrict__ x, float* __restrict__ part, int rows, int C, int G,
                                  int slabs) {
  // grid: (slabs, N), 256 threads = (C/VN channel vectors) x (row lanes).  part layout [N][slabs][G][2].
  // Bit-reproducible: per-thread channel sums go to shared memory and are reduced in a fixed order (the atomics of the
  // first vectorised version made the statistics differ in the last ulp from run to run, and ~40 TF32 layers amplified
  // that to 4e-3 in pixel space -- caught by the batch-independence test).
  // Round-1 v2: 256-row slabs with four 16-byte loads in flight per thread and a two-stage reduction (channel over row
  // lanes by C threads, then group over its channels) -- the 64-row version with a 64-step single-thread tail per group
  // streamed at ~1.9 TB/s (profiles/r01/launches_cfg64_v9.txt).
  constexpr int VN = GnVec<T>::N;
  extern __shared__ float gn_sm[];  // [rlanes][C][2], then [C][2] reduced over the row lanes (aliases lane 0)
  const int n = blockIdx.y, slab = blockIdx.x;
  const int cpg = C / G;
  const int r0 = slab * GN_PART_ROWS;
  const int r1 = min(rows, r0 + GN_PART_ROWS);
  const T* base = x + ((size_t)n * rows) * C;
  const int cvecs = C / VN;
  const int rlanes = blockDim.x / cvecs;          // >= 1 (C/VN <= 256)
  const int cv = threadIdx.x % cvecs, rl = threadIdx.x / cvecs;
  if (rl < rlanes) {
    float s[VN], q[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) { s[i] = 0.f; q[i] = 0.f; }
    int r = r0 + rl;
    for (; r + 3 * rlanes < r1; r += 4 * rlanes) {          // four independent loads, then the (ordered) accumulation
      float f0[VN], f1[VN], f2[VN], f3[VN];
      gn_load<T>(base + (size_t)r * C + cv * VN, f0);
      gn_load<T>(base + (size_t)(r + rlanes) * C + cv * VN, f1);
      gn_load<T>(base + (size_t)(r + 2 * rlanes) * C + cv * VN, f2);
      gn_load<T>(base + (size_t)(r + 3 * rlanes) * C + cv * VN, f3);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        s[i] += f0[i]; q[i] = fmaf(f0[i], f0[i], q[i]);
        s[i] += f1[i]; q[i] = fmaf(f1[i], f1[i], q[i]);
        s[i] += f2[i]; q[i] = fmaf(f2[i], f2[i], q[i]);
        s[i] += f3[i]; q[i] = fmaf(f3[i], f3[i], q[i]);
      }
    }
    for (; r < r1; r += rlanes) {
      float f[VN];
      gn_load<T>(base + (size_t)r * C + cv * VN, f);
#pragma unroll
      for (int i = 0; i < VN; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
    float* dst = gn_sm + ((size_t)rl * C + cv * VN) * 2;
#pragma unroll
    for (int i = 0; i < VN; ++i) { dst[2 * i] = s[i]; dst[2 * i + 1] = q[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {      // stage 1: channel c over the row lanes, fixed order
    float ss = 0.f, qq = 0.f;
    for (int l = 0; l < rlanes; ++l) {
      ss += gn_sm[((size_t)l * C + c) * 2];
      qq += gn_sm[((size_t)l * C + c) * 2 + 1];
    }
    gn_sm[(size_t)c * 2] = ss;                              // lane 0's slot of channel c: only this thread touches it
    gn_sm[(size_t)c * 2 + 1] = qq;
  }
  __syncthreads();
  if ((int)threadIdx.x < G) {                               // stage 2: group over its channels, fixed order
    const int g = threadIdx.x;
    float ss = 0.f, qq = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      ss += gn_sm[(size_t)c * 2];
      qq += gn_sm[(size_t)c * 2 + 1];
    }
    float* o = part + (((size_t)n * slabs + slab) * G + g) * 2;
    o[0] = ss; o[1] = qq;
  }
}

__global__ void gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats, int slabs, int G,
                                   double count, float eps) {
  // grid: N, block: 32 x G threads -- a warp per group strides the slabs, fixed-order shuffle tree, fp64 combine.
  const int n = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g >= G) return;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < slabs; i += 32) {
    const float* p = part + ((size_t)n * slabs + i) * G * 2 + 2 * g;
    s += (double)p[0]; q += (double)p[1];
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if (lane == 0) {
    double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((size_t)n * G + g) * 2] = (float)mean;
    stats[((size_t)n * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// Per-(sample, channel) affine coefficients: y = x * scale + shift with scale = rstd * gamma, shift = beta - mean * scale.
// Folding the statistics lookup out of the streaming kernel leaves it with one FMA (+ SiLU) per element (the first
// vectorised version did 4 dependent __ldg per element and ran at ~50 % of HBM bandwidth on the 256x256 tensors).
__global__ void gn_coeff_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift,
                                int N, int C, int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  const float* st = stats + ((size_t)n * G + c / (C / G)) * 2;
  const float sc = st[1] * gamma[c];
  scale[i] = sc;
  shift[i] = beta[c] - st[0] * sc;
}

int gn_coeff_launch(const float* stats, const float* gamma, const float* beta, float* scale, float* shift, int N, int C, int G,
                    cudaStream_t st) {
  IVG_CHECK(N >= 1 && C >= 1 && G >= 1 && C % G == 0, "groupnorm_coeff: bad shape N=%d C=%d G=%d", N, C, G);
  gn_coeff_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(stats, gamma, beta, scale, shift, N, C, G);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// y = x * scale + shift ; optional SiLU ; optional + pos[(row % pos_rows)][C].
// grid (row chunks, samples); a thread owns ONE 16-byte channel vector (its scale / shift live in registers) and walks the
// chunk's rows with four loads in flight.  (Round-1 v1 was a flat grid-stride loop with two 64-bit divisions and four
// coefficient loads per 16 bytes: ~2.8 TB/s on the decoder tensors.)
constexpr int GN_APPLY_ROWS = 256;
template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float* __restrict__ pos, int rows_per_sample, int C,
                                int silu, int pos_rows) {
  constexpr int VN = GnVec<T>::N;
  const int vecC = C / VN;
  const int rlanes = blockDim.x / vecC;
  const int cv = threadIdx.x % vecC, rl = threadIdx.x / vecC;
  if (rl >= rlanes) return;
  const int n = blockIdx.y;
  const int c = cv * VN;
  float sc[VN], sh[VN];
#pragma unroll
  for (int k = 0; k < VN / 4; ++k) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * C + c) + k);
    const float4 b = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * C + c) + k);
    sc[4 * k] = a.x; sc[4 * k + 1] = a.y; sc[4 * k + 2] = a.z; sc[4 * k + 3] = a.w;
    sh[4 * k] = b.x; sh[4 * k + 1] = b.y; sh[4 * k + 2] = b.z; sh[4 * k + 3] = b.w;
  }
  const int r0 = blockIdx.x * GN_APPLY_ROWS;
  const int r1 = min(rows_per_sample, r0 + GN_APPLY_ROWS);
  const size_t sbase = (size_t)n * rows_per_sample;
  auto finish = [&](float (&f)[VN], int r) {
#pragma unroll
    for (int k = 0; k < VN; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
    if (silu) {
#pragma unroll
      for (int k = 0; k < VN; ++k) f[k] = silu_f(f[k]);
    }
    if (pos) {
      const float4* pr = reinterpret_cast<const float4*>(pos + ((sbase + r) % pos_rows) * C + c);
#pragma unroll
      for (int k = 0; k < VN / 4; ++k) {
        const float4 a = __ldg(pr + k);
        f[4 * k] += a.x; f[4 * k + 1] += a.y; f[4 * k + 2] += a.z; f[4 * k + 3] += a.w;
      }
    }
    gn_store<T>(y + (sbase + r) * C + c, f);
  };
  int r = r0 + rl;
  for (; r + 3 * rlanes < r1; r += 4 * rlanes) {
    float f0[VN], f1[VN], f2[VN], f3[VN];
    gn_load<T>(x + (sbase + r) * C + c, f0);
    gn_load<T>(x + (sbase + r + rlanes) * C + c, f1);
    gn_load<T>(x + (sbase + r + 2 * rlanes) * C + c, f2);
    gn_load<T>(x + (sbase + r + 3 * rlanes) * C + c, f3);
    finish(f0, r); finish(f1, r + rlanes); finish(f2, r + 2 * rlanes); finish(f3, r + 3 * rlanes);
  }
  for (; r < r1; r += rlanes) {
    float f[VN];
    gn_load<T>(x + (sbase + r) * C + c, f);
    finish(f, r);
  }
}

int gn_finalize_launch(const float* part, float* stats, int N, int slabs, int G, double count, float eps,
                       cudaStream_t st) {
  IVG_CHECK(G >= 1 && G <= 32, "groupnorm_finalize: bad G=%d", G);
  if (N <= 0) return 0;
  gn_finalize_kernel<<<N, 32 * G, 0, st>>>(part, stats, slabs, G, count, eps);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int gn_stats_launch_t(const T* x, float* part_ws, float* stats, int N, int rows, int C, int G, float eps,
                      cudaStream_t st) {
  const int slabs = cdiv(rows, GN_PART_ROWS);
  dim3 grid(slabs, N);
  const size_t smem = (size_t)(256 / (C / GnVec<T>::N)) * C * 2 * sizeof(float);
  gn_partial_kernel<T><<<grid, 256, smem, st>>>(x, part_ws, rows, C, G, slabs);
  gn_finalize_kernel<<<N, 32 * G, 0, st>>>(part_ws, stats, slabs, G, (double)rows * (C / G), eps);
  count_launch(2);
  IVG_LAUNCH_CHECK();
  return 0;
}

int gn_stats_launch(int dtype, const void* x, float* part_ws, float* stats, int N, int rows, int C, int G, float eps,
                    cudaStream_t st) {
  IVG_CHECK(C % G == 0 && C % 8 == 0 && C / (dtype == DT_BF16 ? 8 : 4) <= 256 && G <= 32, "groupnorm: bad C=%d G=%d", C, G);
  IVG_CHECK(((uintptr_t)x & 15) == 0, "groupnorm: x must be 16-byte aligned");
  if (N == 0) return 0;
  if (dtype == DT_BF16) return gn_stats_launch_t((const __nv_bfloat16*)x, part_ws, stats, N, rows, C, G, eps, st);
  return gn_stats_launch_t((const float*)x, part_ws, stats, N, rows, C, G, eps, st);
}

int gn_apply_launch(int dtype, const void* x, void* y, const float* stats, const float* gamma, const float* beta,
                    const float* pos, float* coef_ws, long long total_rows, int rows_per_sample, int C, int G, int silu,
                    int pos_rows, cudaStream_t st) {
  if (total_rows == 0) return 0;
  IVG_CHECK(coef_ws != nullptr, "groupnorm_apply: coefficient workspace [2 * samples * C] is required");
  const int N = (int)(total_rows / rows_per_sample);
  float* scale = coef_ws;
  float* shift = coef_ws + (size_t)N * C;
  gn_coeff_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(stats, gamma, beta, scale, shift, N, C, G);
  const int vecC = C / (dtype == DT_BF16 ? 8 : 4);
  IVG_CHECK(vecC >= 1 && vecC <= 256 && C % (dtype == DT_BF16 ? 8 : 4) == 0, "groupnorm_apply: bad C=%d", C);
  IVG_CHECK((long long)N * rows_per_sample == total_rows && N <= 65535, "groupnorm_apply: rows %lld / samples %d", total_rows, N);
  const int threads = (256 / vecC) * vecC;               // (row lanes) x (channel vectors), <= 256
  dim3 grid(cdiv(rows_per_sample, GN_APPLY_ROWS), N);
  if (dtype == DT_BF16)
    gn_apply_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, scale, shift, pos,
                                                             rows_per_sample, C, silu, pos_rows);
  else
    gn_apply_kernel<float><<<grid, threads, 0, st>>>((const float*)x, (float*)y, scale, shift, pos, rows_per_sample, C, silu,
                                                     pos_rows);
  count_launch(2);
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// conv_in: 3x3, pad 1, Cin=3 -> Cout (vae.py:86,149).  Reads the caller's NCHW fp32 pixels, writes NHWC T.
// K = 27 is far too small for a tensor-core tile; the kernel is bound by the output write.
// Thread = (pixel, 8 output channels); weights [Cout][27] + bias staged in shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                               T* __restrict__ y, int N, int H, int W, int Cout, int fpc, int clipT, int foff) {
  extern __shared__ float cw[];  // [Cout*27] + [Cout]
  for (int i = threadIdx.x; i < Cout * 27; i += blockDim.x) cw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) cw[Cout * 27 + i] = b[i];
  __syncthreads();
  const int groups = Cout / 8;
  const long long total = (long long)N * H * W * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long pix = i / groups;
    const int xw = (int)(pix % W);
    const int yh = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    const size_t nf = (size_t)(n / fpc) * clipT + foff + (n % fpc);  // frame slot inside the caller's [B,T,3,H,W]
    float in[27];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = yh + ky - 1, xx = xw + kx - 1;
          in[c * 9 + ky * 3 + kx] =
              (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + ((nf * 3 + c) * H + yy) * W + xx) : 0.f;
        }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wj = cw + (g * 8 + j) * 27;
      float a = cw[Cout * 27 + g * 8 + j];
#pragma unroll
      for (int t = 0; t < 27; ++t) a = fmaf(in[t], wj[t], a);
      o[j] = a;
    }
    T* dst = y + pix * Cout + g * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = from_f32<T>(o[j]);
  }
}

int conv_in_launch(int dtype, const float* x, const float* w, const float* b, void* y, int N, int H, int W, int Cout,
                   int fpc, int clipT, int foff, cudaStream_t st) {
  IVG_CHECK(fpc > 0, "conv_in: frames-per-clip must be positive");
  IVG_CHECK(Cout % 8 == 0, "conv_in: Cout %% 8 != 0");
  if (N == 0) return 0;
  size_t smem = (size_t)Cout * 28 * sizeof(float);
  long long work = (long long)N * H * W * (Cout / 8);
  int blocks = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
  if (dtype == DT_BF16)
    conv_in_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>(x, w, b, (__nv_bfloat16*)y, N, H, W, Cout, fpc, clipT, foff);
  else
    conv_in_kernel<float><<<blocks, 256, smem, st>>>(x, w, b, (float*)y, N, H, W, Cout, fpc, clipT, foff);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// conv_out of the decoders: GroupNorm -> SiLU -> 3x3 conv C -> 3 (vae.py:292-294,363-369), fused.
// One warp per output pixel: lanes split the channels, 9 taps x (C/32) channels each, 3 warp reductions.
// Reads raw NHWC T activations + GN stats, writes the caller-visible NCHW fp32 frame.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void conv_out3_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ w /*[3][9][C]*/, const float* __restrict__ b,
                                 float* __restrict__ y, int N, int H, int W, int C, int G, int fpc, int clipT, int foff) {
  extern __shared__ float ws[];  // [27*C] weights, then gamma[C], beta[C]
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { ws[27 * C + i] = gamma[i]; ws[28 * C + i] = beta[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long total = (long long)N * H * W;
  const int cpg = C / G;
  for (long long pix = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < total;
       pix += (long long)gridDim.x * warps_per_block) {
    const int xw = (int)(pix % W);
    const int yh = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int c = lane * 2; c < C; c += 64) {
      const float* st0 = stats + ((size_t)n * G + c / cpg) * 2;
      const float* st1 = stats + ((size_t)n * G + (c + 1) / cpg) * 2;
      const float m0 = st0[0], r0 = st0[1] * ws[27 * C + c], bb0 = ws[28 * C + c];
      const float m1 = st1[0], r1 = st1[1] * ws[27 * C + c + 1], bb1 = ws[28 * C + c + 1];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = yh + t / 3 - 1, xx = xw + t % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const T* px = x + (((size_t)n * H + yy) * W + xx) * C + c;
        float v0 = silu_f((to_f32(px[0]) - m0) * r0 + bb0);
        float v1 = silu_f((to_f32(px[1]) - m1) * r1 + bb1);
        a0 = fmaf(v0, ws[(0 * 9 + t) * C + c], a0); a0 = fmaf(v1, ws[(0 * 9 + t) * C + c + 1], a0);
        a1 = fmaf(v0, ws[(1 * 9 + t) * C + c], a1); a1 = fmaf(v1, ws[(1 * 9 + t) * C + c + 1], a1);
        a2 = fmaf(v0, ws[(2 * 9 + t) * C + c], a2); a2 = fmaf(v1, ws[(2 * 9 + t) * C + c + 1], a2);
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if (lane == 0) {
      const size_t plane = (size_t)H * W;
      const size_t nf = (size_t)(n / fpc) * clipT + foff + (n % fpc);
      float* dst = y + nf * 3 * plane + (size_t)yh * W + xw;
      dst[0] = a0 + b[0]; dst[plane] = a1 + b[1]; dst[2 * plane] = a2 + b[2];
    }
  }
}

// Tiled variant for C <= 128: a CTA owns an 8 x 16 output tile; the (8+2) x (16+2) halo is normalised + activated
// ONCE into shared memory (fp32, [pixel][C]) instead of once per tap, then each warp produces output pixels with lanes
// splitting the channels.  The first version re-read and re-normalised every input pixel 9 times from L1/L2.
constexpr int CO3_TH = 8, CO3_TW = 16;
template <typename T>
__global__ void __launch_bounds__(256)
conv_out3_tiled_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ w /*[3][9][C]*/,
                       const float* __restrict__ b, float* __restrict__ y, int N, int H, int W, int C, int G, int fpc,
                       int clipT, int foff) {
  constexpr int VN = GnVec<T>::N;
  extern __shared__ float co_sm[];
  float* act = co_sm;                                   // [(TH+2)*(TW+2)][C]
  float* ws = act + (CO3_TH + 2) * (CO3_TW + 2) * C;   // [27][C]
  const int tiles_x = W / CO3_TW, tiles_y = H / CO3_TH;
  const int tile = blockIdx.x;
  const int n = tile / (tiles_x * tiles_y);
  const int t2 = tile - n * tiles_x * tiles_y;
  const int y0 = (t2 / tiles_x) * CO3_TH, x0 = (t2 % tiles_x) * CO3_TW;
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) ws[i] = w[i];
  const int cvecs = C / VN;
  const int halo = (CO3_TH + 2) * (CO3_TW + 2);
  for (int i = threadIdx.x; i < halo * cvecs; i += blockDim.x) {
    const int cv = i % cvecs, hp = i / cvecs;
    const int hy = hp / (CO3_TW + 2), hx = hp % (CO3_TW + 2);
    const int yy = y0 + hy - 1, xx = x0 + hx - 1;
    float f[VN];
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      gn_load<T>(x + (((size_t)n * H + yy) * W + xx) * C + cv * VN, f);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const int c = cv * VN + k;
        const float* st = stats + ((size_t)n * G + c / cpg) * 2;
        f[k] = silu_f((f[k] - __ldg(st)) * __ldg(st + 1) * __ldg(gamma + c) + __ldg(beta + c));
      }
    } else {
#pragma unroll
      for (int k = 0; k < VN; ++k) f[k] = 0.f;      // zero padding applies to the ACTIVATED tensor
    }
#pragma unroll
    for (int k = 0; k < VN; ++k) act[(size_t)hp * C + cv * VN + k] = f[k];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)H * W;
  const size_t nf = (size_t)(n / fpc) * clipT + foff + (n % fpc);
  // each warp produces 4 horizontally adjacent output pixels at a time: the 3 x 6 input window is read once for the four
  // of them and every weight vector once per group (0.1 shared-memory loads per FMA instead of 0.33: the one-pixel version
  // was shared-memory-bandwidth bound at 2.8 ms per launch, profiles/r01/launches_cfg64_v7.txt)
  for (int grp = warp; grp < CO3_TH * CO3_TW / 4; grp += 8) {
    const int py = grp / (CO3_TW / 4), px0 = (grp % (CO3_TW / 4)) * 4;
    float acc[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; }
    for (int c = lane * 4; c < C; c += 128) {
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        float4 v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k)
          v[k] = *reinterpret_cast<const float4*>(act + (size_t)((py + ty) * (CO3_TW + 2) + px0 + k) * C + c);
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
          const int t = ty * 3 + tx;
          const float4 w0 = *reinterpret_cast<const float4*>(ws + (0 * 9 + t) * C + c);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + (1 * 9 + t) * C + c);
          const float4 w2 = *reinterpret_cast<const float4*>(ws + (2 * 9 + t) * C + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 a = v[j + tx];
            acc[j][0] = fmaf(a.x, w0.x, acc[j][0]); acc[j][0] = fmaf(a.y, w0.y, acc[j][0]);
            acc[j][0] = fmaf(a.z, w0.z, acc[j][0]); acc[j][0] = fmaf(a.w, w0.w, acc[j][0]);
            acc[j][1] = fmaf(a.x, w1.x, acc[j][1]); acc[j][1] = fmaf(a.y, w1.y, acc[j][1]);
            acc[j][1] = fmaf(a.z, w1.z, acc[j][1]); acc[j][1] = fmaf(a.w, w1.w, acc[j][1]);
            acc[j][2] = fmaf(a.x, w2.x, acc[j][2]); acc[j][2] = fmaf(a.y, w2.y, acc[j][2]);
            acc[j][2] = fmaf(a.z, w2.z, acc[j][2]); acc[j][2] = fmaf(a.w, w2.w, acc[j][2]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc[j][o] += __shfl_xor_sync(0xffffffffu, acc[j][o], off);
    if (lane < 12) {                       // lane = 4 * o + j writes output channel o of pixel j
      const int o = lane >> 2, j = lane & 3;
      float r = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int oo = 0; oo < 3; ++oo) r = (jj == j && oo == o) ? acc[jj][oo] : r;
      y[nf * 3 * plane + (size_t)o * plane + (size_t)(y0 + py) * W + x0 + px0 + j] = r + __ldg(b + o);
    }
  }
}

int conv_out3_launch(int dtype, const void* x, const float* stats, const float* gamma, const float* beta,
                     const float* w, const float* b, float* y, int N, int H, int W, int C, int G, int fpc, int clipT,
                     int foff, cudaStream_t st) {
  IVG_CHECK(fpc > 0, "conv_out3: frames-per-clip must be positive");
  IVG_CHECK(C % 2 == 0 && C % G == 0, "conv_out3: bad C");
  if (N == 0) return 0;
  if (C <= 128 && C % 4 == 0 && C % (dtype == DT_BF16 ? 8 : 4) == 0 && H % CO3_TH == 0 && W % CO3_TW == 0) {
    const size_t tsm = (size_t)((CO3_TH + 2) * (CO3_TW + 2) * C + 27 * C) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.pending()) {
      IVG_CUDA(cudaFuncSetAttribute(conv_out3_tiled_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      IVG_CUDA(cudaFuncSetAttribute(conv_out3_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      attr_once.mark();
    }
    const long long tiles = (long long)N * (H / CO3_TH) * (W / CO3_TW);
    IVG_CHECK(tiles < 0x7fffffffLL && tsm <= 110 * 1024, "conv_out3: tile count / shared memory out of range");
    if (dtype == DT_BF16)
      conv_out3_tiled_kernel<__nv_bfloat16><<<(int)tiles, 256, tsm, st>>>((const __nv_bfloat16*)x, stats, gamma, beta, w, b,
                                                                         y, N, H, W, C, G, fpc, clipT, foff);
    else
      conv_out3_tiled_kernel<float><<<(int)tiles, 256, tsm, st>>>((const float*)x, stats, gamma, beta, w, b, y, N, H, W, C, G,
                                                                 fpc, clipT, foff);
    count_launch();
    IVG_LAUNCH_CHECK();
    return 0;
  }
  size_t smem = (size_t)29 * C * sizeof(float);
  long long pixels = (long long)N * H * W;
  int blocks = (int)((pixels + 7) / 8 < 148 * 8 ? (pixels + 7) / 8 : 148 * 8);
  if (dtype == DT_BF16)
    conv_out3_kernel<__nv_bfloat16><<<blocks, 256, smem, st>>>((const __nv_bfloat16*)x, stats, gamma, beta, w, b, y,
                                                               N, H, W, C, G, fpc, clipT, foff);
  else
    conv_out3_kernel<float><<<blocks, 256, smem, st>>>((const float*)x, stats, gamma, beta, w, b, y, N, H, W, C, G, fpc, clipT, foff);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// nearest-neighbour 2x upsample, NHWC (diffusers Upsample2D -> F.interpolate(scale_factor=2, 'nearest'))
// ---------------------------------------------------------------------------------------------
template <typename V>
__global__ void upsample2x_kernel(const V* __restrict__ x, V* __restrict__ y, int N, int H, int W, int CV) {
  const long long total = (long long)N * (2 * H) * (2 * W) * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV);
    long long p = i / CV;
    const int xo = (int)(p % (2 * W)); p /= (2 * W);
    const int yo = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    y[i] = x[(((size_t)n * H + (yo >> 1)) * W + (xo >> 1)) * CV + c];
  }
}

int upsample2x_launch(int dtype, const void* x, void* y, int N, int H, int W, int C, cudaStream_t st) {
  const int esz = dtype == DT_BF16 ? 2 : 4;
  IVG_CHECK((C * esz) % 16 == 0, "upsample2x: C*elsize must be a multiple of 16 bytes");
  if (N == 0) return 0;
  const int CV = C * esz / 16;
  long long work = (long long)N * 4 * H * W * CV;
  int blocks = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  upsample2x_kernel<uint4><<<blocks, 256, 0, st>>>((const uint4*)x, (uint4*)y, N, H, W, CV);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// patchify (compressive_vq_model.py:192-195): NHWC [F,16,16,C] -> [F*16, p*p*C], feature order (py,px,c);
// depatchify (:247-250, einsum nhwpqc->nchpwq) is the inverse.  p = 4.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ x, T* __restrict__ y, int F, int R, int C, int P, int inverse) {
  const int PR = R / P;  // patches per side
  const long long total = (long long)F * R * R * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int xx = (int)(p % R); p /= R;
    const int yy = (int)(p % R);
    const int f = (int)(p / R);
    const long long patch_row = ((long long)f * PR + yy / P) * PR + xx / P;
    const long long j = patch_row * ((long long)P * P * C) + ((yy % P) * P + (xx % P)) * C + c;
    if (inverse) y[i] = x[j]; else y[j] = x[i];
  }
}

int patchify_launch(int dtype, const void* x, void* y, int F, int R, int C, int P, int inverse, cudaStream_t st) {
  if (F == 0) return 0;
  long long work = (long long)F * R * R * C;
  int blocks = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  if (dtype == DT_BF16)
    patchify_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, F, R, C, P, inverse);
  else
    patchify_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, F, R, C, P, inverse);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// dtype conversion (fp32 <-> T), used at API boundaries (latents to fp32 for the exact VQ stage)
// ---------------------------------------------------------------------------------------------
template <typename A, typename B>
__global__ void convert_kernel(const A* __restrict__ x, B* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f32<B>(to_f32(x[i]));
}

int convert_launch(int src_dtype, const void* x, int dst_dtype, void* y, long long n, cudaStream_t st) {
  if (n == 0) return 0;
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (src_dtype == DT_F32 && dst_dtype == DT_BF16)
    convert_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)y, n);
  else if (src_dtype == DT_BF16 && dst_dtype == DT_F32)
    convert_kernel<__nv_bfloat16, float><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)y, n);
  else if (src_dtype == DT_F32 && dst_dtype == DT_F32)
    convert_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, n);
  else
    convert_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Token (de)serialisation, compressive_vq_model.py:205-220 and :227-236.  int64 throughout.
// Sequence layout for t context frames / f future frames (cr = 256 ctx tokens, dr = 16 dyn tokens):
//   c0[256] scf c1[256] ... | sdf d0[16] sdf d1[16] ...      length L = t*257 - 1 + f*17
// ---------------------------------------------------------------------------------------------
__global__ void serialise_kernel(const long long* __restrict__ ic, const long long* __restrict__ id,
                                 long long* __restrict__ tokens, long long* __restrict__ labels, int B, int t, int f,
                                 int cr, int dr, long long nvq, long long ndyn) {
  const int L = t * (cr + 1) - 1 + f * (dr + 1);
  const long long total = (long long)B * L;
  const long long scf = nvq + ndyn, sdf = nvq + ndyn + 1;
  const int ctx_len = t * (cr + 1) - 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / L), pos = (int)(i % L);
    long long tok, lab;
    if (pos < ctx_len) {
      const int fr = pos / (cr + 1), k = pos % (cr + 1);
      tok = (k == cr) ? scf : ic[((size_t)b * t + fr) * cr + k];
      lab = -100;
    } else {
      const int q = pos - ctx_len;
      const int fr = q / (dr + 1), k = q % (dr + 1);
      tok = (k == 0) ? sdf : id[((size_t)b * f + fr) * dr + (k - 1)] + nvq;
      lab = (q == 0) ? -100 : tok;
    }
    tokens[i] = tok;
    if (labels) labels[i] = lab;
  }
}

int serialise_launch(const long long* ic, const long long* id, long long* tokens, long long* labels, int B, int t,
                     int f, int cr, int dr, long long nvq, long long ndyn, cudaStream_t st) {
  if (B == 0) return 0;
  const long long total = (long long)B * (t * (cr + 1) - 1 + f * (dr + 1));
  serialise_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(ic, id, tokens, labels, B, t, f, cr, dr, nvq, ndyn);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// tokens[B, L] -> gathered codebook rows: ctx latents [B*t*cr, D] and dyn latents [B*f*dr, D] (T)
template <typename T>
__global__ void detok_gather_kernel(const long long* __restrict__ tokens, const float* __restrict__ cb_ctx,
                                    const float* __restrict__ cb_dyn, T* __restrict__ qc, T* __restrict__ qd, int B,
                                    int t, int f, int cr, int dr, int D, long long nvq, long long ndyn, int L,
                                    int* __restrict__ bad_ctx) {
  const long long nctx = (long long)B * t * cr, ndy = (long long)B * f * dr;
  const int ctx_len = t * (cr + 1) - 1;
  const long long total = (nctx + ndy) * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long row = i / D;
    if (row < nctx) {
      const int b = (int)(row / ((long long)t * cr));
      const int r = (int)(row % ((long long)t * cr));
      const int fr = r / cr, k = r % cr;
      long long id = tokens[(size_t)b * L + fr * (cr + 1) + k];
      if (id < 0 || id >= nvq) {       // the reference's embedding lookup raises here (:238): flag it, the host raises
        if (bad_ctx != nullptr && d == 0) *bad_ctx = 1;
        id = id < 0 ? 0 : nvq - 1;     // stay in bounds; the output of this call is discarded by the caller
      }
      qc[row * D + d] = from_f32<T>(cb_ctx[id * D + d]);
    } else {
      const long long r2 = row - nctx;
      const int b = (int)(r2 / ((long long)f * dr));
      const int r = (int)(r2 % ((long long)f * dr));
      const int fr = r / dr, k = r % dr;
      long long id = tokens[(size_t)b * L + ctx_len + fr * (dr + 1) + 1 + k] - nvq;
      id = id < 0 ? 0 : (id > ndyn - 1 ? ndyn - 1 : id);  // clamp, compressive_vq_model.py:236
      qd[r2 * D + d] = from_f32<T>(cb_dyn[id * D + d]);
    }
  }
}

int detok_gather_launch(int dtype, const long long* tokens, const float* cb_ctx, const float* cb_dyn, void* qc,
                        void* qd, int B, int t, int f, int cr, int dr, int D, long long nvq, long long ndyn, int L,
                        int* bad_ctx, cudaStream_t st) {
  if (B == 0) return 0;
  const long long total = ((long long)B * t * cr + (long long)B * f * dr) * D;
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == DT_BF16)
    detok_gather_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(tokens, cb_ctx, cb_dyn, (__nv_bfloat16*)qc,
                                                               (__nv_bfloat16*)qd, B, t, f, cr, dr, D, nvq, ndyn, L, bad_ctx);
  else
    detok_gather_kernel<float><<<blocks, 256, 0, st>>>(tokens, cb_ctx, cb_dyn, (float*)qc, (float*)qd, B, t, f, cr, dr,
                                                       D, nvq, ndyn, L, bad_ctx);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Input pipeline (reference inference/utils.py:12-16 NPZParser.preprocess, ivideogpt/data/simple_dataloader.py:394,510):
//     images = images / 255 ; images = torchvision resize(images, [S, S])       (bilinear, antialias, align_corners=False)
// One thread per output pixel, all channels; the anti-aliased triangle filter of ATen's _upsample_bilinear2d_aa is
// evaluated in place (weights normalised per axis, horizontal pass inside the vertical one, sequential mul + add like the
// CPU kernel, no FMA contraction).  Input: uint8 or fp32 frames with arbitrary element strides ([T,H,W,C] as stored in the
// .npz episodes, or the [T,C,H,W] float tensor the reference's preprocess() is handed); output fp32 [T, C, OH, OW].
// ---------------------------------------------------------------------------------------------
constexpr int RS_MAXTAPS = 40;      // 2 * ceil(scale) + 2 for down-scaling factors up to 19
struct RsAxis { int lo, n; float w[RS_MAXTAPS]; };

__device__ __forceinline__ void rs_axis(int in_size, int out_size, int i, RsAxis& a) {
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float inv = scale >= 1.0f ? __fdiv_rn(1.0f, scale) : 1.0f;
  const float center = __fmul_rn(scale, (float)i + 0.5f);
  int lo = (int)(__fadd_rn(__fsub_rn(center, support), 0.5f));
  lo = lo < 0 ? 0 : lo;
  int hi = (int)(__fadd_rn(__fadd_rn(center, support), 0.5f));
  hi = hi > in_size ? in_size : hi;
  int n = hi - lo;
  n = n > RS_MAXTAPS ? RS_MAXTAPS : n;
  float total = 0.f;
  for (int j = 0; j < n; ++j) {
    float x = __fmul_rn(__fadd_rn(__fsub_rn(__fadd_rn((float)j, (float)lo), center), 0.5f), inv);
    x = fabsf(x);
    const float w = x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f;
    a.w[j] = w;
    total = __fadd_rn(total, w);
  }
  for (int j = 0; j < n; ++j) a.w[j] = __fdiv_rn(a.w[j], total);
  a.lo = lo; a.n = n;
}

template <typename TIN>
__global__ void resize_aa_kernel(const TIN* __restrict__ in, long long st, long long sy, long long sx, long long sc, int T,
                                 int H, int W, int C, float* __restrict__ out, int OH, int OW, float divisor) {
  const long long total = (long long)T * OH * OW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW);
    const int oy = (int)((idx / OW) % OH);
    const int t = (int)(idx / ((long long)OW * OH));
    RsAxis ax, ay;
    rs_axis(W, OW, ox, ax);
    rs_axis(H, OH, oy, ay);
    for (int c = 0; c < C; ++c) {
      const TIN* base = in + (long long)t * st + (long long)c * sc;
      float acc = 0.f;
      for (int j = 0; j < ay.n; ++j) {
        const TIN* row = base + (long long)(ay.lo + j) * sy + (long long)ax.lo * sx;
        float h = __fmul_rn(__fdiv_rn((float)row[0], divisor), ax.w[0]);
        for (int i = 1; i < ax.n; ++i) h = __fadd_rn(h, __fmul_rn(__fdiv_rn((float)row[(long long)i * sx], divisor), ax.w[i]));
        acc = j == 0 ? __fmul_rn(h, ay.w[0]) : __fadd_rn(acc, __fmul_rn(h, ay.w[j]));
      }
      out[(((long long)t * C + c) * OH + oy) * OW + ox] = acc;
    }
  }
}

int resize_aa_launch(int in_dtype, const void* in, long long st, long long sy, long long sx, long long sc, int T, int H, int W,
                     int C, float* out, int OH, int OW, float divisor, cudaStream_t stream) {
  IVG_CHECK(in_dtype == DT_F32 || in_dtype == 2, "resize_aa: input dtype %d (0 = fp32, 2 = uint8)", in_dtype);
  IVG_CHECK(T >= 0 && H >= 1 && W >= 1 && C >= 1 && OH >= 1 && OW >= 1 && divisor != 0.f, "resize_aa: bad geometry");
  IVG_CHECK(2 * ((H + OH - 1) / OH) + 2 <= RS_MAXTAPS && 2 * ((W + OW - 1) / OW) + 2 <= RS_MAXTAPS,
            "resize_aa: down-scaling factor above %d is not supported", (RS_MAXTAPS - 2) / 2);
  if (T == 0) return 0;
  const long long total = (long long)T * OH * OW;
  const int blocks = (int)((total + 127) / 128 < 148 * 16 ? (total + 127) / 128 : 148 * 16);
  if (in_dtype == 2)
    resize_aa_kernel<unsigned char><<<blocks, 128, 0, stream>>>((const unsigned char*)in, st, sy, sx, sc, T, H, W, C, out, OH, OW, divisor);
  else
    resize_aa_kernel<float><<<blocks, 128, 0, stream>>>((const float*)in, st, sy, sx, sc, T, H, W, C, out, OH, OW, divisor);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// VectorQuantizer.forward beyond the argmin (diffusers VectorQuantizer(beta, legacy=False); reference call sites
// compressive_vq_model.py:297-301 inside decode(), reached from forward() :332-369 -- tokenizer training / evaluation):
//     z_q = E[idx]                                   (returned in the decoder's compute dtype; straight-through value)
//     loss = beta * mean((sg(z_q) - z)^2) + mean((z_q - sg(z))^2)  =  (beta + 1) * mean((z_q - z)^2)   in value.
// One pass: gather + squared difference, fixed-size grid of per-block partial sums, then a single-block ordered sum
// (deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int VQC_BLOCKS = 296;
template <typename T>
__global__ void vq_commit_partial_kernel(const float* __restrict__ z, const float* __restrict__ cb, const long long* __restrict__ idx,
                                         T* __restrict__ zq, long long N, int D, long long K, float* __restrict__ part) {
  __shared__ float s_red[8];
  float acc = 0.f;
  const long long total = N * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / D;
    const int d = (int)(i - n * D);
    long long id = idx[n];
    id = id < 0 ? 0 : (id >= K ? K - 1 : id);
    const float q = cb[id * D + d];
    const float df = q - z[i];
    acc = fmaf(df, df, acc);
    if (zq != nullptr) zq[i] = from_f32<T>(q);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    part[blockIdx.x] = t;
  }
}
__global__ void vq_commit_final_kernel(const float* __restrict__ part, int nblocks, double scale, float* __restrict__ loss) {
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nblocks; ++i) t += (double)part[i];
    *loss = (float)(t * scale);
  }
}
int vq_commit_launch(int dtype, const float* z, const float* cb, const long long* idx, void* zq, long long N, int D, long long K,
                     float beta, float* part_ws, float* loss, cudaStream_t st) {
  IVG_CHECK(N >= 1 && D >= 1 && K >= 1, "vq_commit: bad shape N=%lld D=%d K=%lld", N, D, K);
  if (dtype == DT_BF16)
    vq_commit_partial_kernel<__nv_bfloat16><<<VQC_BLOCKS, 256, 0, st>>>(z, cb, idx, (__nv_bfloat16*)zq, N, D, K, part_ws);
  else
    vq_commit_partial_kernel<float><<<VQC_BLOCKS, 256, 0, st>>>(z, cb, idx, (float*)zq, N, D, K, part_ws);
  vq_commit_final_kernel<<<1, 32, 0, st>>>(part_ws, VQC_BLOCKS, (double)(beta + 1.0f) / ((double)N * (double)D), loss);
  count_launch(2);
  IVG_LAUNCH_CHECK();
  return 0;
}

}  // namespace ivg
