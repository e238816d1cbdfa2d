// Parameter blocks shared by gemm_tc.cu (device) and capi.cu (host descriptor translation).
#pragma once
#include "common.cuh"

namespace ivg {

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 224;      // warp 0 TMA producer, warp 1 + warp 6 tcgen05.mma issuers, warps 2-5 epilogue
#ifndef IVG_XF_WARPS
#define IVG_XF_WARPS 8
#endif
constexpr int GEMM_XF_WARPS = IVG_XF_WARPS;                  // operand-transform warps (4, 8 or 16)
constexpr int GEMM_THREADS_XF = GEMM_THREADS + 32 * GEMM_XF_WARPS;    // GroupNorm + SiLU applied to the A tiles in shared memory
constexpr int GEMM_ROWB = 128;  // bytes of K per k-block row (one 128B swizzle atom)

struct alignas(64) GemmMaps {
  CUtensorMap a[5];  // mode 0: a[0] (3-D: K, M, batch).  mode 1: a[0..3] (4-D: C, W, H, N) + a[4] extra source
  CUtensorMap b;     // 3-D: K, N, batch
};

struct GemmParams {
  int M, N, num_kb;
  int mode;
  // ---- mode 1 geometry ----
  int H, W, tw, th;  // output spatial size and tile shape (tw*th == 128)
  int cpb;           // channel chunks (k-blocks) per tap
  int ntaps;
  int extra_kb;      // k-blocks taken from a[4] (after ntaps*cpb)
  signed char tap_map[9], tap_dy[9], tap_dx[9];
  // ---- batching / heads (mode 0) ----
  int batch;         // grid z extent (= outer * heads)
  int heads;
  int a_bsel, a_bdiv, b_bsel, b_bdiv, o_bsel;  // *_bsel: 0 -> 0, 1 -> outer / div, 2 -> bz
  int a_kbase, a_khead, b_kbase, b_khead, b_nhead, o_nhead;  // element offsets (per head)
  int causal_skip;   // skip tiles entirely above the diagonal (n0 > m0 + 127)
  // ---- epilogue ----
  void* out;
  long long ldo, out_bstride;
  int out_dtype;     // DT_F32 / DT_BF16
  const float* bias;
  int bias_along_m;
  const void* residual;
  long long ldr, res_bstride;
  int res_dtype;
  int act;           // 0 none, 1 SiLU, 2 SwiGLU over interleaved column pairs (out has N/2 columns)
  float alpha;
  int tiles_m, tiles_n;
  int issuers;       // 1 or 2 tcgen05.mma issuing warps (2: alternate k-blocks, same accumulator; see gemm_tc.cu)
  int mh2;           // allow 256 x 256 CTA tiles (BN = 256 launches with enough tiles; opt-in with IVGPT_GEMM_MH2=1: measured slower than 128 x 256)
  // ---- fused GroupNorm statistics of the OUTPUT (mode 1): per (image, tile, n-tile, epilogue warp) partial sums ----
  float* gn_part;    // [images][slabs][gn_groups][2] (sum, sum of squares), slabs = tiles_per_image * tiles_n * 4; or null
  int gn_groups;
  // ---- fused GroupNorm (+ SiLU) of the INPUT (mode 1, stride 1): y = silu(x * xf_scale[img][c] + xf_shift[img][c]) is applied
  //      to every A tile after it lands in shared memory and before the tensor core reads it; zero padding stays zero ----
  const float* xf_scale;   // [images][Cin] or null
  const float* xf_shift;
  int xf_silu;
  int xf_cin;
};

}  // namespace ivg
