// Backward pass of the compressive tokenizer (reference train_tokenizer.py:734 `accelerator.backward(loss)` through
// compressive_vq_model.py:332-369 / vae.py / conditional_vae.py): the memory-bound pieces.  Every contraction of the
// backward (conv dgrad = the forward conv kernel on flipped weights, conv wgrad = dY^T x im2col(X)^T, Linear / attention
// products) runs on the tcgen05 GEMM / conv kernel of gemm_tc.cu; this file holds what surrounds them, all fp32 NHWC:
//   colsum            bias gradients                           out[c] (+)= sum_m x[m][c]
//   groupnorm_bwd     GroupNorm (+ SiLU) backward              4 small kernels, fixed summation order
//   transpose_pad     K-major operands of the 3x3 weight gradient: channel-major copies with a zero frame per image
//   im2col3x3_t       the same for stride-2 / 3-channel convs   colT[tap*C + c][pixel]  (transposed while gathering)
//   zero_insert2x     operand of the stride-2 conv dgrad       out[2i+1][2j+1] = dy[i][j], zeros elsewhere
//   upsample2x_bwd    nearest-neighbour 2x backward            sum of the 4 children
//   silu / silu_bwd, axpby, reduce_mid (sum over a middle axis), nchw <-> nhwc with channel padding, vq_bwd
// These kernels move O(bytes of the activation) each; they are written for coalesced channel-fastest access, grid-stride
// loops sized against the SM count, and deterministic two-stage reductions (no float atomics).
#include "../../include/ivgpt_b200.h"
#include "common.cuh"

namespace ivg {
namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  return (int)(g > cap ? cap : g);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) {
  const float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}

// ---------------------------------------------------------------------------------------------
// column sums
// ---------------------------------------------------------------------------------------------
__global__ void colsum_part_kernel(const float* __restrict__ x, long long M, int C, long long ld, float* __restrict__ part) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (long long m = blockIdx.x; m < M; m += gridDim.x) s += x[m * ld + c];
  part[(long long)blockIdx.x * C + c] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int nb, int C, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += part[(long long)b * C + c];
  out[c] = accumulate ? out[c] + s : s;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm (+ SiLU) backward.  y = act(gamma * xhat + beta), xhat = (x - mean) * rstd per (sample, group).
//   d0 = dy * act'(gamma*xhat+beta);  dgamma = sum d0*xhat;  dbeta = sum d0
//   dx = rstd * (gamma*d0 - (A + xhat*Bq)/m),  A = sum_group gamma*d0,  Bq = sum_group gamma*d0*xhat,  m = rows * C/G
// ---------------------------------------------------------------------------------------------
__global__ void gn_bwd_part_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int silu, int rows, int C,
                                   int G, int chunks, float* __restrict__ part /* [S][chunks][C][2] */) {
  const int s = blockIdx.y, ch = blockIdx.x;
  const int r0 = (int)((long long)rows * ch / chunks), r1 = (int)((long long)rows * (ch + 1) / chunks);
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = stats[((long long)s * G + g) * 2], rstd = stats[((long long)s * G + g) * 2 + 1];
    const float ga = gamma[c], be = beta[c];
    const float* xp = x + ((long long)s * rows + r0) * C + c;
    const float* dp = dy + ((long long)s * rows + r0) * C + c;
    float s1 = 0.f, s2 = 0.f;
    for (int r = r0; r < r1; ++r, xp += C, dp += C) {
      const float xh = (*xp - mean) * rstd;
      float d = *dp;
      if (silu) d *= silu_grad(fmaf(ga, xh, be));
      s1 += d;
      s2 = fmaf(d, xh, s2);
    }
    float* o = part + (((long long)s * chunks + ch) * C + c) * 2;
    o[0] = s1;
    o[1] = s2;
  }
}
// one block per sample: totals per channel, then per group (A, Bq)
__global__ void gn_bwd_group_kernel(const float* __restrict__ part, const float* __restrict__ gamma, int C, int G, int chunks,
                                    float* __restrict__ tot /* [S][C][2] */, float* __restrict__ ab /* [S][G][2] */) {
  extern __shared__ float sh[];  // [C][2]: gamma-weighted totals
  const int s = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float* p = part + (((long long)s * chunks + k) * C + c) * 2;
      t1 += p[0];
      t2 += p[1];
    }
    tot[((long long)s * C + c) * 2] = t1;
    tot[((long long)s * C + c) * 2 + 1] = t2;
    sh[2 * c] = gamma[c] * t1;
    sh[2 * c + 1] = gamma[c] * t2;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < cpg; ++j) {
      a += sh[2 * (g * cpg + j)];
      b += sh[2 * (g * cpg + j) + 1];
    }
    ab[((long long)s * G + g) * 2] = a;
    ab[((long long)s * G + g) * 2 + 1] = b;
  }
}
__global__ void gn_bwd_param_kernel(const float* __restrict__ tot, int S_, int C, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float g = 0.f, b = 0.f;
  for (int s = 0; s < S_; ++s) {
    b += tot[((long long)s * C + c) * 2];
    g += tot[((long long)s * C + c) * 2 + 1];
  }
  dgamma[c] = accumulate ? dgamma[c] + g : g;
  dbeta[c] = accumulate ? dbeta[c] + b : b;
}
__global__ void gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ ab, int silu, int rows, int C, int G, long long total,
                                    float* __restrict__ dx, int accumulate) {
  const int cpg = C / G;
  const float inv_m = 1.f / ((float)rows * (float)cpg);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long s = i / ((long long)rows * C);
    const int g = c / cpg;
    const float mean = stats[(s * G + g) * 2], rstd = stats[(s * G + g) * 2 + 1];
    const float A = ab[(s * G + g) * 2], Bq = ab[(s * G + g) * 2 + 1];
    const float ga = gamma[c];
    const float xh = (x[i] - mean) * rstd;
    float d = dy[i];
    if (silu) d *= silu_grad(fmaf(ga, xh, beta[c]));
    const float v = rstd * (ga * d - (A + xh * Bq) * inv_m);
    dx[i] = accumulate ? dx[i] + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// im2col, transposed: colT[k][p], k = tap*C + c (tap = 3*a + b), p = (n*Ho + y)*Wo + xo.
//   stride 1: source pixel (y + a - 1, xo + b - 1) (zero outside);  stride 2: (2y + a, 2xo + b) (zero when >= H / W:
//   the reference pads bottom/right only, vae.py Downsample2D with padding=0 after F.pad(0,1,0,1)).
// Rows k in [9C, Krows) are written as zeros (row-count padding for 16-byte pitches of the GEMM output).
// ---------------------------------------------------------------------------------------------
__global__ void im2col_t_kernel(const float* __restrict__ x, float* __restrict__ colT, int H, int W, int C, int stride, int Ho,
                                int Wo, long long P, int Krows) {
  __shared__ float tile[32][33];
  const long long p0 = (long long)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  const int k = k0 + threadIdx.x;
  const int tap = k / C, c = k - tap * C;
  const int a = tap / 3, b = tap - 3 * a;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    float v = 0.f;
    if (p < P && tap < 9) {
      const int xo = (int)(p % Wo);
      const long long t = p / Wo;
      const int y = (int)(t % Ho);
      const long long n = t / Ho;
      const int yi = stride == 1 ? y + a - 1 : 2 * y + a;
      const int xi = stride == 1 ? xo + b - 1 : 2 * xo + b;
      if (yi >= 0 && yi < H && xi >= 0 && xi < W) v = x[((n * H + yi) * W + xi) * C + c];
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int kk = k0 + j;
    const long long p = p0 + threadIdx.x;
    if (kk < Krows && p < P) colT[(long long)kk * P + p] = tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------------------------------------
// NHWC -> channel-major with a zero frame around every image:
//   out[c][n*img_stride + (y+1)*Wp + (x+1) - shift] = x[n][y][x][c]        (Wp >= W+2, a multiple of 4)
// (the caller zero-fills `out`; this kernel writes the interior only).  In this layout the input pixel that tap (a, b) of a
// 3x3 / stride-1 / pad-1 conv pairs with output pixel q sits at q + (a-1)*Wp + (b-1): the weight gradient becomes nine
// GEMMs over the pixel axis whose B operand is the same array read at a constant K offset, and image borders need no
// masking because dY is zero on the frame.  A TMA box must start on a 16-byte boundary of the innermost dimension
// (measured: an odd element offset is an illegal instruction), so the row term (a-1)*Wp is a box coordinate and the
// column term b-1 is baked into three copies written with shift = b-1.  Replaces the 9x im2col copy by 3x + 1x.
// ---------------------------------------------------------------------------------------------
__global__ void transpose_pad_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W, int C, int Wp, int shift,
                                     long long img_stride, long long ld_out, int xtiles) {
  __shared__ float tile[32][33];
  const int xt = blockIdx.x % xtiles;
  const long long ny = blockIdx.x / xtiles;       // n*H + y
  const int y = (int)(ny % H);
  const long long n = ny / H;
  const int x0 = xt * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int xx = x0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (xx < W && c < C) ? x[((n * H + y) * W + xx) * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, xx = x0 + threadIdx.x;
    if (c < C && xx < W) out[(long long)c * ld_out + n * img_stride + (long long)(y + 1) * Wp + xx + 1 - shift] = tile[threadIdx.x][j];
  }
}

__global__ void zero_insert2x_kernel(const float* __restrict__ dy, float* __restrict__ out, int h, int w, int C, long long total) {
  // out [N, 2h, 2w, C]
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int xx = (int)(t % (2 * w));
    t /= 2 * w;
    const int yy = (int)(t % (2 * h));
    const long long n = t / (2 * h);
    out[i] = ((xx & 1) && (yy & 1)) ? dy[((n * h + (yy >> 1)) * w + (xx >> 1)) * C + c] : 0.f;
  }
}
__global__ void upsample2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int C, long long total) {
  // dy [N, 2H, 2W, C] -> dx [N, H, W, C]
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int xx = (int)(t % W);
    t /= W;
    const int yy = (int)(t % H);
    const long long n = t / H;
    const float* p = dy + ((n * 2 * H + 2 * yy) * 2 * W + 2 * xx) * C + c;
    const long long rowp = (long long)2 * W * C;
    dx[i] = (p[0] + p[C]) + (p[rowp] + p[rowp + C]);
  }
}
__global__ void silu_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    out[i] = dy ? dy[i] * silu_grad(v) : v * sigmoidf_(v);
  }
}
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, float a, float b,
                             long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}
// out[o][i] (+)= sum_m x[o][m][i]
__global__ void reduce_mid_kernel(const float* __restrict__ x, float* __restrict__ out, long long outer, int mid, long long inner,
                                  int accumulate) {
  const long long total = outer * inner;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
    const long long o = j / inner, i = j - o * inner;
    const float* p = x + o * mid * inner + i;
    float s = 0.f;
    for (int m = 0; m < mid; ++m) s += p[(long long)m * inner];
    out[j] = accumulate ? out[j] + s : s;
  }
}
// y [N,H,W,Cd] <- x [N,Cs,H,W] (channels >= Cs are zero)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int Cs, int Cd, long long HW, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cd);
    const long long t = i / Cd;
    const long long hw = t % HW, n = t / HW;
    y[i] = c < Cs ? x[(n * Cs + c) * HW + hw] : 0.f;
  }
}
// straight-through + commitment gradients of the VQ layer (diffusers VectorQuantizer, legacy=False):
//   loss = beta * mean((sg[zq] - z)^2) + mean((zq - sg[z])^2);   out = z + sg[zq - z]
//   dz = dout + gl * beta * 2/n * (z - zq);   de_rows = gl * 2/n * (zq - z)   (scattered onto the codebook by embed_bwd)
__global__ void vq_bwd_kernel(const float* __restrict__ z, const float* __restrict__ zq, const float* __restrict__ dout,
                              const float* __restrict__ gl, float beta, long long n, float* __restrict__ dz,
                              float* __restrict__ de) {
  const float g = gl ? *gl : 0.f;
  const float k = 2.f * g / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = z[i] - zq[i];
    dz[i] = (dout ? dout[i] : 0.f) + beta * k * d;
    de[i] = -k * d;
  }
}

}  // namespace
}  // namespace ivg

using namespace ivg;

extern "C" {

int ivgpt_colsum(const float* x, long long M, int C, long long ld, float* part_ws, int part_rows, float* out, int accumulate,
                 void* stream) {
  IVG_CHECK(M >= 0 && C > 0 && part_rows > 0, "colsum: bad shape");
  int nb = (int)(M < part_rows ? (M > 0 ? M : 1) : part_rows);
  dim3 g1(nb, (C + 127) / 128);
  colsum_part_kernel<<<g1, 128, 0, S(stream)>>>(x, M, C, ld, part_ws);
  colsum_final_kernel<<<(C + 127) / 128, 128, 0, S(stream)>>>(part_ws, nb, C, out, accumulate);
  count_launch(2);
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_groupnorm_bwd_chunks(int samples, int rows) {
  int chunks = (148 * 4 + samples - 1) / (samples > 0 ? samples : 1);
  if (chunks > rows / 8) chunks = rows / 8;
  if (chunks < 1) chunks = 1;
  if (chunks > 256) chunks = 256;
  return chunks;
}

int ivgpt_groupnorm_bwd(const float* x, const float* dy, const float* stats, const float* gamma, const float* beta, int silu,
                        int samples, int rows, int C, int G, float* ws /* samples*(chunks+1)*C*2 + samples*G*2 floats */,
                        float* dx, int dx_accumulate, float* dgamma, float* dbeta, int param_accumulate, void* stream) {
  IVG_CHECK(samples > 0 && rows > 0 && C > 0 && G > 0 && C % G == 0, "groupnorm_bwd: bad shape (S=%d rows=%d C=%d G=%d)", samples,
            rows, C, G);
  IVG_CHECK(C * 2 * sizeof(float) <= 48 * 1024, "groupnorm_bwd: C=%d too large", C);
  const int chunks = ivgpt_groupnorm_bwd_chunks(samples, rows);
  float* part = ws;
  float* tot = part + (long long)samples * chunks * C * 2;
  float* ab = tot + (long long)samples * C * 2;
  cudaStream_t st = S(stream);
  gn_bwd_part_kernel<<<dim3(chunks, samples), 256, 0, st>>>(x, dy, stats, gamma, beta, silu, rows, C, G, chunks, part);
  gn_bwd_group_kernel<<<samples, 256, C * 2 * sizeof(float), st>>>(part, gamma, C, G, chunks, tot, ab);
  int n = 2;
  if (dgamma) {
    gn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, st>>>(tot, samples, C, dgamma, dbeta, param_accumulate);
    ++n;
  }
  if (dx) {
    const long long total = (long long)samples * rows * C;
    gn_bwd_apply_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, dy, stats, gamma, beta, ab, silu, rows, C, G, total, dx,
                                                               dx_accumulate);
    ++n;
  }
  count_launch(n);
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_im2col3x3_t(const float* x, float* colT, int N, int H, int W, int C, int stride, int k_rows, void* stream) {
  IVG_CHECK(stride == 1 || stride == 2, "im2col3x3_t: stride %d", stride);
  IVG_CHECK(k_rows >= 9 * C, "im2col3x3_t: k_rows %d < 9*C", k_rows);
  const int Ho = H / stride, Wo = W / stride;
  const long long P = (long long)N * Ho * Wo;
  if (P <= 0) return 0;
  IVG_CHECK((P + 31) / 32 < 2147483647LL, "im2col3x3_t: too many pixels");
  dim3 grid((unsigned)((P + 31) / 32), (k_rows + 31) / 32);
  im2col_t_kernel<<<grid, dim3(32, 8), 0, S(stream)>>>(x, colT, H, W, C, stride, Ho, Wo, P, k_rows);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_transpose_pad(const float* x, float* out, int N, int H, int W, int C, int Wp, int shift, long long img_stride,
                        long long ld_out, void* stream) {
  IVG_CHECK(Wp >= W + 2 && Wp % 4 == 0 && shift >= -1 && shift <= 1, "transpose_pad: Wp=%d (W=%d) shift=%d", Wp, W, shift);
  IVG_CHECK(img_stride >= (long long)(H + 2) * Wp + 4 && ld_out >= N * img_stride, "transpose_pad: strides too small");
  if (N <= 0) return 0;
  const int xtiles = (W + 31) / 32;
  const long long gx = (long long)N * H * xtiles;
  IVG_CHECK(gx < 2147483647LL, "transpose_pad: too many rows");
  transpose_pad_kernel<<<dim3((unsigned)gx, (C + 31) / 32), dim3(32, 8), 0, S(stream)>>>(x, out, H, W, C, Wp, shift, img_stride,
                                                                                          ld_out, xtiles);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_zero_insert2x(const float* dy, float* out, int N, int h, int w, int C, void* stream) {
  const long long total = (long long)N * 4 * h * w * C;
  if (total <= 0) return 0;
  zero_insert2x_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(dy, out, h, w, C, total);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_upsample2x_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
  const long long total = (long long)N * H * W * C;
  if (total <= 0) return 0;
  upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(dy, dx, H, W, C, total);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_silu(const float* x, const float* dy, float* out, long long n, void* stream) {
  if (n <= 0) return 0;
  silu_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(x, dy, out, n);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_axpby(const float* x, const float* y, float* out, float a, float b, long long n, void* stream) {
  if (n <= 0) return 0;
  axpby_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(x, y, out, a, b, n);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_reduce_mid(const float* x, float* out, long long outer, int mid, long long inner, int accumulate, void* stream) {
  if (outer * inner <= 0) return 0;
  reduce_mid_kernel<<<grid_for(outer * inner, 256), 256, 0, S(stream)>>>(x, out, outer, mid, inner, accumulate);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_nchw_to_nhwc(const float* x, float* y, long long N, int Cs, int Cd, long long HW, void* stream) {
  IVG_CHECK(Cs > 0 && Cd >= Cs, "nchw_to_nhwc: Cs=%d Cd=%d", Cs, Cd);
  const long long total = N * HW * Cd;
  if (total <= 0) return 0;
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, S(stream)>>>(x, y, Cs, Cd, HW, total);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

int ivgpt_vq_bwd(const float* z, const float* zq, const float* dout, const float* gloss, float beta, long long n, float* dz,
                 float* de_rows, void* stream) {
  if (n <= 0) return 0;
  vq_bwd_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(z, zq, dout, gloss, beta, n, dz, de_rows);
  count_launch();
  IVG_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
