// extern "C" entry points of libivgpt_b200.so (declared in include/ivgpt_b200.h) + host-side plumbing:
// thread-local error string, tensor-map construction through the driver entry point, GEMM/conv descriptors.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/ivgpt_b200.h"
#include "common.cuh"
#include "gemm_params.cuh"
#include "decode_mega.cuh"

namespace ivg {

static thread_local char g_err[1024] = "";
unsigned long long g_launches = 0;
bool g_pdl = false;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---- tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tensor_map(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  IVG_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  IVG_CHECK(((uintptr_t)base & 15) == 0, "tensor map: base pointer %p not 16-byte aligned", base);
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) {
    gs[i] = strides_bytes[i];
    IVG_CHECK(gs[i] % 16 == 0, "tensor map: stride %d = %llu bytes is not a multiple of 16", i,
              (unsigned long long)gs[i]);
  }
  CUresult r = fn(out, dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IVG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)",
            (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
            (unsigned long long)(rank > 2 ? gd[2] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0);
  return 0;
}

static int g_num_sms[64] = {0};      // per device ordinal (a process may drive several GPUs)
static int num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = g_num_sms[dev & 63];
  if (!n) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  return n;
}

// tcgen05.mma issuing warps of the GEMM / conv kernel: 1 (default) or 2 (IVGPT_MMA_ISSUERS=2 / ivgpt_set_deterministic(0) after
// opting in).  Two issuers reach the full tensor rate in isolation (tools/probes/mma_probe.cu) but changed nothing in the real
// kernel, whose limit is operand supply from L2 (profiles/r02/gemm_issuers_and_tiles_ab.txt), and their summation order is not
// fixed; together with the 256 x 256 CTA tiles they faulted on one launch shape, so those tiles always run with one issuer.
static int g_issuers_optin = -1;
static int g_issuers = 0;
static int mma_issuers() {
  if (g_issuers_optin < 0) {
    const char* e = getenv("IVGPT_MMA_ISSUERS");
    g_issuers_optin = (e != nullptr && e[0] == '2') ? 1 : 0;
    const char* d = getenv("IVGPT_DETERMINISTIC");
    g_issuers = (g_issuers_optin && !(d != nullptr && d[0] == '1')) ? 2 : 1;
  }
  return g_issuers;
}
static int g_mh2 = -1;
static int gemm_mh2() {
  if (g_mh2 < 0) {
    const char* e = getenv("IVGPT_GEMM_MH2");
    g_mh2 = (e != nullptr && e[0] == '1') ? 1 : 0;      // opt-in: measured slower (profiles/r02/gemm_issuers_and_tiles_ab.txt)
  }
  return g_mh2;
}

// ---- forward declarations of launchers defined in the other translation units -----------------
int gemm_tc_dispatch(int dtype, int bn, const GemmMaps& maps, const GemmParams& p, int num_sms, cudaStream_t stream);
void gemm_profile_enable(int on);
int gemm_profile_collect(int bucket, double* ms_total, double* flops_total, long long* launches);
int profile_begin(int bucket, double work, cudaStream_t stream);
void profile_end(int idx, cudaStream_t stream);
int vq_argmin_launch(const float*, const float*, float*, unsigned long long*, long long*, int, int, int, int,
                     cudaStream_t);
extern int g_vq_order;
int gn_stats_launch(int, const void*, float*, float*, int, int, int, int, float, cudaStream_t);
int gn_finalize_launch(const float*, float*, int, int, int, double, float, cudaStream_t);
int gn_apply_launch(int, const void*, void*, const float*, const float*, const float*, const float*, float*, long long,
                    int, int, int, int, int, cudaStream_t);
int conv_in_launch(int, const float*, const float*, const float*, void*, int, int, int, int, int, int, int,
                   cudaStream_t);
int conv_out3_launch(int, const void*, const float*, const float*, const float*, const float*, const float*, float*,
                     int, int, int, int, int, int, int, int, cudaStream_t);
int upsample2x_launch(int, const void*, void*, int, int, int, int, cudaStream_t);
int patchify_launch(int, const void*, void*, int, int, int, int, int, cudaStream_t);
int gn_coeff_launch(const float*, const float*, const float*, float*, float*, int, int, int, cudaStream_t);
int vq_commit_launch(int, const float*, const float*, const long long*, void*, long long, int, long long, float, float*, float*,
                     cudaStream_t);
int convert_launch(int, const void*, int, void*, long long, cudaStream_t);
int serialise_launch(const long long*, const long long*, long long*, long long*, int, int, int, int, int, long long,
                     long long, cudaStream_t);
int detok_gather_launch(int, const long long*, const float*, const float*, void*, void*, int, int, int, int, int, int,
                        long long, long long, int, int*, cudaStream_t);
int embed_launch(const long long*, long long, int, const int*, const float*, float*, long long, int, long long,
                 cudaStream_t);
int add_rows_launch(float*, const float*, long long, cudaStream_t);
int rmsnorm_launch(int, const float*, const float*, void*, long long, int, float, cudaStream_t);
int rope_kv_launch(int, const void*, void*, void*, void*, void*, int, int, int, int, int, const int*, const float*,
                   const float*, cudaStream_t);
int softmax_launch(int, const float*, void*, long long, int, int, long long, long long, int, int, cudaStream_t);
int decode_attn_launch(int, const void*, const void*, const void*, void*, int, int, int, int, const int*, float,
                       cudaStream_t);
int argmax_launch(const float*, long long, int, int, long long*, long long, const int*, cudaStream_t);
int topk_sample_launch(const float*, long long, int, int, int, float, unsigned long long, unsigned long long,
                       long long*, long long, const int*, const unsigned long long*, cudaStream_t);
int ce_loss_launch(const float*, long long, int, int, int, const long long*, float*, float*, float*, cudaStream_t);
int incr_launch(int*, int, cudaStream_t);
int resize_aa_launch(int, const void*, long long, long long, long long, long long, int, int, int, int, float*, int, int, float,
                     cudaStream_t);
int slot_embed_add_launch(float*, const float*, const int*, int, int, int, int, int, cudaStream_t);
int slot_force_launch(long long*, long long, const int*, int, int, int, long long, cudaStream_t);
int transpose_launch(int, const void*, void*, int, int, int, long long, long long, long long, long long, cudaStream_t);
int swiglu_launch(int, int, const void*, const void*, void*, long long, cudaStream_t);
int rmsnorm_bwd_launch(int, const float*, const float*, const void*, float*, float*, float*, long long, int, float,
                       cudaStream_t);
int softmax_bwd_launch(int, const void*, const float*, void*, long long, int, int, long long, int, float, cudaStream_t);
int rope_bwd_launch(int, const float*, const float*, const float*, void*, int, int, int, const float*, const float*,
                    cudaStream_t);
int ce_bwd_launch(int, const float*, long long, int, int, int, const long long*, const float*, float, void*, long long,
                  cudaStream_t);
int embed_bwd_launch(const long long*, const float*, float*, long long, int, long long, cudaStream_t);
int adamw_launch(float*, const float*, float*, float*, long long, float, float, float, float, float, int, float,
                 cudaStream_t);
int add_to_f32_launch(int, float*, const void*, long long, cudaStream_t);
int dropout_launch(int, const void*, void*, long long, float, unsigned long long, cudaStream_t);
int flash_attn_launch(const void*, const void*, const void*, void*, float*, int, int, int, int, long long, long long, long long,
                      long long, long long, int, float, int, cudaStream_t);
int decode_attn_fused_launch(int, const void*, void*, void*, void*, int, int, int, int, const int*, const float*,
                             const float*, float, cudaStream_t);

}  // namespace ivg


using namespace ivg;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int esize(int dt) { return dt == DT_BF16 ? 2 : 4; }

extern "C" {

const char* ivgpt_last_error(void) { return ivg::last_error(); }
unsigned long long ivgpt_launch_count(void) { return ivg::g_launches; }

int ivgpt_profile_enable(int on) { gemm_profile_enable(on); return 0; }
int ivgpt_profile_collect(int bucket, double* ms_total, double* flops_total, long long* launches) {
  return gemm_profile_collect(bucket, ms_total, flops_total, launches);
}
int ivgpt_count_add(long long n) { ivg::g_launches += (unsigned long long)n; return 0; }

int ivgpt_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  IVG_CUDA(cudaGetDevice(&dev));
  IVG_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  IVG_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  IVG_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}

int ivgpt_vq_argmin(const float* z, const float* codebook, float* enorm_ws, unsigned long long* packed_ws,
                    long long* idx, int N, int K, int D, void* stream) {
  return vq_argmin_launch(z, codebook, enorm_ws, packed_ws, idx, N, K, D, num_sms(), S(stream));
}

static int pick_bn(int N, long long tiles_m_times_batch) {
  // widest tile (<= 128) that still gives every SM a CTA; skinny problems (decode steps) get narrow tiles so
  // that more CTAs stream the weight matrix concurrently.
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  const int sms = num_sms();
  // 256-wide tiles cut operand traffic per FLOP by 25 % (A 16 KB + B 32 KB per 128x256x64 MACs = 96 B/clk/SM instead
  // of 128): the 128x128 kernel was measured L2->SM bound (tensor pipe 31-47 %, profiles/r01), so take them whenever
  // they divide N and still leave >= 2 tiles per SM.
  if (N % 256 == 0 && tiles_m_times_batch * (N / 256) >= 2LL * sms) return 256;
  const int cands[3] = {128, 64, 32};
  for (int i = 0; i < 3; ++i)
    if (tiles_m_times_batch * ((N + cands[i] - 1) / cands[i]) >= sms) return cands[i];
  return 32;
}

int ivgpt_gemm(const ivgpt_gemm_desc* d, void* stream) {
  IVG_CHECK(d != nullptr, "gemm: null descriptor");
  IVG_CHECK(d->dtype == DT_F32 || d->dtype == DT_BF16, "gemm: bad dtype %d", d->dtype);
  if (d->M <= 0 || d->N <= 0 || d->batch <= 0) return 0;
  IVG_CHECK(d->K > 0, "gemm: K must be positive");
  const int es = esize(d->dtype);
  const int BK = 128 / es;
  const int heads = d->heads > 0 ? d->heads : 1;
  IVG_CHECK(d->batch % heads == 0, "gemm: batch %d not a multiple of heads %d", d->batch, heads);
  IVG_CHECK(heads == 1 || d->K % BK == 0 || (d->a_khead == 0 && d->b_khead == 0),
            "gemm: per-head K offsets need K %% %d == 0 (K=%d)", BK, d->K);
  // a TMA box must start on a 16-byte boundary of the innermost (K) dimension: an odd element offset is not a wrong answer
  // but an illegal instruction on sm_100a (measured), so it is refused here
  IVG_CHECK((d->a_kbase * es) % 16 == 0 && (d->a_khead * es) % 16 == 0 && (d->b_kbase * es) % 16 == 0 && (d->b_khead * es) % 16 == 0,
            "gemm: K offsets (a_kbase=%d a_khead=%d b_kbase=%d b_khead=%d) must be multiples of %d elements", d->a_kbase,
            d->a_khead, d->b_kbase, d->b_khead, 16 / es);
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const long long tiles_m = (d->M + GEMM_BM - 1) / GEMM_BM;
  int bn = d->bn ? d->bn : pick_bn(d->N, tiles_m * d->batch);
  IVG_CHECK(bn == 32 || bn == 64 || bn == 128 || bn == 256, "gemm: bad bn %d", bn);
  {
    uint64_t dims[3] = {(uint64_t)d->a_cols, (uint64_t)d->a_rows, (uint64_t)(d->a_batches > 0 ? d->a_batches : 1)};
    uint64_t str[2] = {(uint64_t)d->lda * es, (uint64_t)(d->a_batches > 1 ? d->a_bstride : (long long)d->a_rows * d->lda) * es};
    uint32_t box[3] = {(uint32_t)BK, GEMM_BM, 1};
    if (make_tensor_map(&maps.a[0], d->dtype, d->a, 3, dims, str, box, 1)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->b_cols, (uint64_t)d->b_rows, (uint64_t)(d->b_batches > 0 ? d->b_batches : 1)};
    uint64_t str[2] = {(uint64_t)d->ldb * es, (uint64_t)(d->b_batches > 1 ? d->b_bstride : (long long)d->b_rows * d->ldb) * es};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)bn, 1};
    if (make_tensor_map(&maps.b, d->dtype, d->b, 3, dims, str, box, 1)) return 1;
  }
  p.M = d->M; p.N = d->N; p.num_kb = (d->K + BK - 1) / BK;
  p.mode = 0;
  p.batch = d->batch; p.heads = heads;
  p.a_bsel = d->a_bsel; p.a_bdiv = d->a_bdiv > 0 ? d->a_bdiv : 1;
  p.b_bsel = d->b_bsel; p.b_bdiv = d->b_bdiv > 0 ? d->b_bdiv : 1;
  p.o_bsel = d->o_bsel;
  p.a_kbase = d->a_kbase; p.a_khead = d->a_khead; p.b_kbase = d->b_kbase; p.b_khead = d->b_khead;
  p.b_nhead = d->b_nhead; p.o_nhead = d->o_nhead;
  p.causal_skip = d->causal_skip;
  p.out = d->out; p.ldo = d->ldo; p.out_bstride = d->out_bstride; p.out_dtype = d->out_dtype;
  p.bias = d->bias; p.bias_along_m = d->bias_along_m;
  p.residual = d->residual; p.ldr = d->ldr; p.res_bstride = d->res_bstride; p.res_dtype = d->res_dtype;
  p.act = d->act; p.alpha = d->alpha;
  p.tiles_m = (int)tiles_m; p.tiles_n = (d->N + bn - 1) / bn;
  IVG_CHECK(d->act != IVGPT_ACT_SWIGLU || (d->N % 2 == 0 && d->residual == nullptr), "gemm: SwiGLU needs even N, no residual");
  p.issuers = mma_issuers(); p.mh2 = gemm_mh2();
  return gemm_tc_dispatch(d->dtype, bn, maps, p, num_sms(), S(stream));
}

int ivgpt_conv3x3_plan(const ivgpt_conv_desc* d, int* bn_out, int* gn_slabs) {
  IVG_CHECK(d != nullptr && bn_out != nullptr && gn_slabs != nullptr, "conv3x3_plan: null argument");
  const int Hout = d->Hin / d->stride, Wout = d->Win / d->stride;
  const int tw = Wout < 128 ? Wout : 128;
  IVG_CHECK(tw > 0 && 128 % tw == 0, "conv3x3_plan: output width %d unsupported", Wout);
  const int th = 128 / tw;
  IVG_CHECK(Hout % th == 0 && Wout % tw == 0, "conv3x3_plan: output %dx%d not tileable", Hout, Wout);
  const long long tiles_img = (long long)(Hout / th) * (Wout / tw);
  const int bn = d->bn ? d->bn : pick_bn(d->Cout, (long long)d->N * tiles_img);
  *bn_out = bn;
  *gn_slabs = (int)(tiles_img * ((d->Cout + bn - 1) / bn) * 4);
  return 0;
}

int ivgpt_groupnorm_finalize(const float* part, float* stats, int samples, int slabs, int G, double count, float eps,
                             void* stream) {
  return gn_finalize_launch(part, stats, samples, slabs, G, count, eps, S(stream));
}

int ivgpt_conv3x3(const ivgpt_conv_desc* d, void* stream) {
  IVG_CHECK(d != nullptr, "conv3x3: null descriptor");
  IVG_CHECK(d->dtype == DT_F32 || d->dtype == DT_BF16, "conv3x3: bad dtype %d", d->dtype);
  IVG_CHECK(d->stride == 1 || d->stride == 2, "conv3x3: stride must be 1 or 2");
  if (d->N <= 0) return 0;
  const int es = esize(d->dtype);
  const int BK = 128 / es;
  IVG_CHECK(d->Cin % BK == 0 && d->C2 % BK == 0, "conv3x3: Cin=%d / C2=%d must be multiples of %d", d->Cin, d->C2, BK);
  const int Hout = d->Hin / d->stride, Wout = d->Win / d->stride;
  IVG_CHECK(d->Hin % d->stride == 0 && d->Win % d->stride == 0, "conv3x3: odd input size for stride 2");
  const int tw = Wout < 128 ? Wout : 128;
  IVG_CHECK(tw > 0 && 128 % tw == 0, "conv3x3: output width %d must divide 128 or be a multiple of it", Wout);
  const int th = 128 / tw;
  IVG_CHECK(Wout % tw == 0 && Hout % th == 0, "conv3x3: output %dx%d not tileable by %dx%d", Hout, Wout, th, tw);
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const uint32_t box[4] = {(uint32_t)BK, (uint32_t)tw, (uint32_t)th, 1};
  const long long C = d->Cin;
  if (d->stride == 1) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)d->Win, (uint64_t)d->Hin, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)C * es, (uint64_t)d->Win * C * es, (uint64_t)d->Hin * d->Win * C * es};
    if (make_tensor_map(&maps.a[0], d->dtype, d->x, 4, dims, str, box, 1)) return 1;
    for (int t = 0; t < 9; ++t) { p.tap_map[t] = 0; p.tap_dy[t] = (signed char)(t / 3 - 1); p.tap_dx[t] = (signed char)(t % 3 - 1); }
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const char* base = reinterpret_cast<const char*>(d->x) + ((long long)py * d->Win + px) * C * es;
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)(d->Win / 2), (uint64_t)(d->Hin / 2), (uint64_t)d->N};
        uint64_t str[3] = {(uint64_t)2 * C * es, (uint64_t)2 * d->Win * C * es, (uint64_t)d->Hin * d->Win * C * es};
        if (make_tensor_map(&maps.a[py * 2 + px], d->dtype, base, 4, dims, str, box, 1)) return 1;
      }
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      p.tap_map[t] = (signed char)((ky & 1) * 2 + (kx & 1));
      p.tap_dy[t] = (signed char)(ky >> 1);
      p.tap_dx[t] = (signed char)(kx >> 1);
    }
  }
  if (d->C2 > 0) {
    IVG_CHECK(d->x2 != nullptr, "conv3x3: C2 > 0 but x2 is null");
    uint64_t dims[4] = {(uint64_t)d->C2, (uint64_t)Wout, (uint64_t)Hout, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)d->C2 * es, (uint64_t)Wout * d->C2 * es, (uint64_t)Hout * Wout * d->C2 * es};
    if (make_tensor_map(&maps.a[4], d->dtype, d->x2, 4, dims, str, box, 1)) return 1;
  }
  const long long Ktot = 9LL * C + d->C2;
  const long long tiles_m = (long long)d->N * (Hout / th) * (Wout / tw);
  int bn = d->bn ? d->bn : pick_bn(d->Cout, tiles_m);
  IVG_CHECK(bn == 32 || bn == 64 || bn == 128 || bn == 256, "conv3x3: bad bn %d", bn);
  {
    uint64_t dims[3] = {(uint64_t)Ktot, (uint64_t)d->Cout, 1};
    uint64_t str[2] = {(uint64_t)Ktot * es, (uint64_t)Ktot * d->Cout * es};
    uint32_t bbox[3] = {(uint32_t)BK, (uint32_t)bn, 1};
    if (make_tensor_map(&maps.b, d->dtype, d->w, 3, dims, str, bbox, 1)) return 1;
  }
  p.M = (int)((long long)d->N * Hout * Wout);
  IVG_CHECK((long long)d->N * Hout * Wout < 0x7fffffffLL, "conv3x3: too many output pixels for one launch");
  p.N = d->Cout; p.num_kb = (int)(Ktot / BK);
  p.mode = 1; p.H = Hout; p.W = Wout; p.tw = tw; p.th = th; p.cpb = d->Cin / BK; p.ntaps = 9; p.extra_kb = d->C2 / BK;
  p.batch = 1; p.heads = 1; p.a_bdiv = 1; p.b_bdiv = 1;
  p.out = d->out; p.ldo = d->Cout; p.out_dtype = d->out_dtype;
  p.bias = d->bias;
  p.residual = d->residual; p.ldr = d->Cout; p.res_dtype = d->res_dtype;
  p.act = d->act; p.alpha = 1.0f;
  p.tiles_m = (int)tiles_m; p.tiles_n = (d->Cout + bn - 1) / bn;
  if (d->gn_part) {
    IVG_CHECK(d->gn_groups > 0 && d->gn_groups <= 32 && d->Cout % d->gn_groups == 0, "conv3x3: bad gn_groups %d", d->gn_groups);
    p.gn_part = d->gn_part; p.gn_groups = d->gn_groups;
  }
  if (d->in_scale) {
    IVG_CHECK(d->in_shift != nullptr && d->stride == 1, "conv3x3: fused input GroupNorm needs in_shift and stride 1");
    p.xf_scale = d->in_scale; p.xf_shift = d->in_shift; p.xf_silu = d->in_silu; p.xf_cin = d->Cin;
  }
  p.issuers = mma_issuers(); p.mh2 = gemm_mh2();
  return gemm_tc_dispatch(d->dtype, bn, maps, p, num_sms(), S(stream));
}

int ivgpt_groupnorm_coeff(const float* stats, const float* gamma, const float* beta, float* scale, float* shift, int samples,
                          int C, int G, void* stream) {
  return gn_coeff_launch(stats, gamma, beta, scale, shift, samples, C, G, S(stream));
}

int ivgpt_groupnorm_stats(int dtype, const void* x, float* part_ws, float* stats, int N, int rows, int C, int G,
                          float eps, void* stream) {
  return gn_stats_launch(dtype, x, part_ws, stats, N, rows, C, G, eps, S(stream));
}
int ivgpt_groupnorm_apply(int dtype, const void* x, void* y, const float* stats, const float* gamma,
                          const float* beta, const float* pos, float* coef_ws, long long total_rows,
                          int rows_per_sample, int C, int G, int silu, int pos_rows, void* stream) {
  return gn_apply_launch(dtype, x, y, stats, gamma, beta, pos, coef_ws, total_rows, rows_per_sample, C, G, silu,
                         pos_rows, S(stream));
}
int ivgpt_conv_in(int dtype, const float* x, const float* w, const float* b, void* y, int N, int H, int W, int Cout,
                  int frames_per_clip, int clip_frames, int frame_offset, void* stream) {
  return conv_in_launch(dtype, x, w, b, y, N, H, W, Cout, frames_per_clip, clip_frames, frame_offset, S(stream));
}
int ivgpt_conv_out3(int dtype, const void* x, const float* stats, const float* gamma, const float* beta,
                    const float* w, const float* b, float* y, int N, int H, int W, int C, int G, int frames_per_clip,
                    int clip_frames, int frame_offset, void* stream) {
  return conv_out3_launch(dtype, x, stats, gamma, beta, w, b, y, N, H, W, C, G, frames_per_clip, clip_frames,
                          frame_offset, S(stream));
}
int ivgpt_upsample2x(int dtype, const void* x, void* y, int N, int H, int W, int C, void* stream) {
  return upsample2x_launch(dtype, x, y, N, H, W, C, S(stream));
}
int ivgpt_patchify(int dtype, const void* x, void* y, int F, int R, int C, int P, int inverse, void* stream) {
  return patchify_launch(dtype, x, y, F, R, C, P, inverse, S(stream));
}
int ivgpt_vq_commit(int dtype, const float* z, const float* codebook, const long long* idx, void* zq, long long N, int D,
                    long long K, float beta, float* part_ws, float* loss, void* stream) {
  return vq_commit_launch(dtype, z, codebook, idx, zq, N, D, K, beta, part_ws, loss, S(stream));
}
int ivgpt_convert(int src_dtype, const void* x, int dst_dtype, void* y, long long n, void* stream) {
  return convert_launch(src_dtype, x, dst_dtype, y, n, S(stream));
}
int ivgpt_tokens_serialise(const long long* ic, const long long* id, long long* tokens, long long* labels, int B,
                           int t, int f, int cr, int dr, long long n_vq, long long n_dyn, void* stream) {
  return serialise_launch(ic, id, tokens, labels, B, t, f, cr, dr, n_vq, n_dyn, S(stream));
}
int ivgpt_tokens_gather(int dtype, const long long* tokens, const float* cb_ctx, const float* cb_dyn, void* qc,
                        void* qd, int B, int t, int f, int cr, int dr, int D, long long n_vq, long long n_dyn, int L,
                        int* bad_ctx, void* stream) {
  return detok_gather_launch(dtype, tokens, cb_ctx, cb_dyn, qc, qd, B, t, f, cr, dr, D, n_vq, n_dyn, L, bad_ctx, S(stream));
}
int ivgpt_embed(const long long* ids, long long ids_stride, int L, const int* dpos, const float* table, float* x,
                long long M, int hidden, long long vocab, void* stream) {
  return embed_launch(ids, ids_stride, L, dpos, table, x, M, hidden, vocab, S(stream));
}
int ivgpt_add_rows(float* x, const float* e, long long n, void* stream) { return add_rows_launch(x, e, n, S(stream)); }
int ivgpt_rmsnorm(int dtype, const float* x, const float* w, void* y, long long M, int hidden, float eps,
                  void* stream) {
  return rmsnorm_launch(dtype, x, w, y, M, hidden, eps, S(stream));
}
int ivgpt_rope_kv(int dtype, const void* qkv, void* q_out, void* k_cache, void* v_cache_t, void* v_rows, int B, int Lq,
                  int heads, int Lmax, int pos0, const int* dpos, const float* cos_tab, const float* sin_tab,
                  void* stream) {
  return rope_kv_launch(dtype, qkv, q_out, k_cache, v_cache_t, v_rows, B, Lq, heads, Lmax, pos0, dpos, cos_tab,
                        sin_tab, S(stream));
}
int ivgpt_softmax(int dtype, const float* Sm, void* P, long long rows, int Lq, int Lk, long long lds, long long ldp,
                  int causal, int causal_off, void* stream) {
  return softmax_launch(dtype, Sm, P, rows, Lq, Lk, lds, ldp, causal, causal_off, S(stream));
}
int ivgpt_decode_attn(int dtype, const void* q, const void* k_cache, const void* v_cache_t, void* out, int B,
                      int heads, int Lmax, int Lcur, const int* dpos, float scale, void* stream) {
  return decode_attn_launch(dtype, q, k_cache, v_cache_t, out, B, heads, Lmax, Lcur, dpos, scale, S(stream));
}
int ivgpt_argmax(const float* logits, long long ld, int rows, int V, long long* out, long long out_stride,
                 const int* dpos, void* stream) {
  return argmax_launch(logits, ld, rows, V, out, out_stride, dpos, S(stream));
}
int ivgpt_topk_sample(const float* logits, long long ld, int rows, int V, int k, float temperature,
                      unsigned long long seed, unsigned long long step, long long* out, long long out_stride,
                      const int* dpos, const unsigned long long* dseed, void* stream) {
  return topk_sample_launch(logits, ld, rows, V, k, temperature, seed, step, out, out_stride, dpos, dseed, S(stream));
}
int ivgpt_ce_loss(const float* logits, long long ld, int B, int L, int V, const long long* labels, float* loss_rows,
                  float* valid_ws, float* loss_out, void* stream) {
  return ce_loss_launch(logits, ld, B, L, V, labels, loss_rows, valid_ws, loss_out, S(stream));
}
int ivgpt_incr(int* p, int by, void* stream) { return incr_launch(p, by, S(stream)); }
int ivgpt_preprocess_resize(int in_dtype, const void* frames, long long stride_t, long long stride_y, long long stride_x,
                            long long stride_c, int T, int H, int W, int C, float* out, int out_h, int out_w, float divisor,
                            void* stream) {
  return resize_aa_launch(in_dtype, frames, stride_t, stride_y, stride_x, stride_c, T, H, W, C, out, out_h, out_w, divisor,
                          S(stream));
}
int ivgpt_slot_embed_add(float* x, const float* slot_emb, const int* dpos, int B, int hidden, int slot0, int period,
                         int nslots, void* stream) {
  return slot_embed_add_launch(x, slot_emb, dpos, B, hidden, slot0, period, nslots, S(stream));
}
int ivgpt_slot_force(long long* tokens, long long tok_stride, const int* dpos, int B, int slot0, int period,
                     long long token, void* stream) {
  return slot_force_launch(tokens, tok_stride, dpos, B, slot0, period, token, S(stream));
}
int ivgpt_decode_attn_fused(int dtype, const void* qkv, void* k_cache, void* v_cache_t, void* out, int B, int heads,
                            int Lmax, int pos, const int* dpos, const float* cos_tab, const float* sin_tab,
                            float scale, void* stream) {
  return decode_attn_fused_launch(dtype, qkv, k_cache, v_cache_t, out, B, heads, Lmax, pos, dpos, cos_tab, sin_tab, scale,
                                  S(stream));
}
int ivgpt_set_pdl(int on) { ivg::g_pdl = on != 0; return 0; }

// ---- training (backward) pieces ------------------------------------------------------------------------
int ivgpt_transpose(int dtype, const void* in, void* out, int batch, int rows, int cols, long long ld_in,
                    long long ld_out, long long bs_in, long long bs_out, void* stream) {
  return transpose_launch(dtype, in, out, batch, rows, cols, ld_in, ld_out, bs_in, bs_out, S(stream));
}
int ivgpt_swiglu(int dtype, int backward, const void* gu, const void* dact, void* out, long long n, void* stream) {
  return swiglu_launch(dtype, backward, gu, dact, out, n, S(stream));
}
int ivgpt_rmsnorm_bwd(int dtype, const float* x, const float* w, const void* dy, float* dres, float* dw_part, float* dw,
                      long long M, int hidden, float eps, void* stream) {
  return rmsnorm_bwd_launch(dtype, x, w, dy, dres, dw_part, dw, M, hidden, eps, S(stream));
}
int ivgpt_softmax_bwd(int dtype, const void* P, const float* dP, void* dS, long long rows, int Lq, int Lk, long long ld,
                      int causal, float scale, void* stream) {
  return softmax_bwd_launch(dtype, P, dP, dS, rows, Lq, Lk, ld, causal, scale, S(stream));
}
int ivgpt_rope_bwd(int dtype, const float* dq, const float* dk, const float* dv, void* dqkv, int B, int L, int heads,
                   const float* cos_tab, const float* sin_tab, void* stream) {
  return rope_bwd_launch(dtype, dq, dk, dv, dqkv, B, L, heads, cos_tab, sin_tab, S(stream));
}
int ivgpt_ce_bwd(int dtype, const float* logits, long long ld, int B, int L, int V, const long long* labels,
                 const float* count, float gscale, void* dlogits, long long ldd, void* stream) {
  return ce_bwd_launch(dtype, logits, ld, B, L, V, labels, count, gscale, dlogits, ldd, S(stream));
}
int ivgpt_embed_bwd(const long long* ids, const float* dx, float* dE, long long M, int hidden, long long vocab,
                    void* stream) {
  return embed_bwd_launch(ids, dx, dE, M, hidden, vocab, S(stream));
}
int ivgpt_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, float gscale, void* stream) {
  return adamw_launch(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, gscale, S(stream));
}
int ivgpt_add_to_f32(int dtype, float* y, const void* x, long long n, void* stream) {
  return add_to_f32_launch(dtype, y, x, n, S(stream));
}
int ivgpt_dropout(int dtype, const void* x, void* y, long long n, float p, unsigned long long seed, void* stream) {
  return dropout_launch(dtype, x, y, n, p, seed, S(stream));
}
int ivgpt_vq_set_order(int order) {
  IVG_CHECK(order == 0 || order == 1, "vq order must be 0 or 1");
  ivg::g_vq_order = order;
  return 0;
}
int ivgpt_vq_get_order(void) { return ivg::g_vq_order; }

// ---- persistent decode megakernel -------------------------------------------------------------------
int ivgpt_mega_layer_bytes(void) { return (int)sizeof(ivg::MegaLayer); }

long long ivgpt_mega_packed_elems(int rows, int cols) {
  return (long long)((rows + ivg::MEGA_BN - 1) / ivg::MEGA_BN) * ivg::MEGA_BN * cols;
}

int ivgpt_mega_pack_weight(const void* w, void* out, int rows, int cols, void* stream) {
  return ivg::mega_pack_weight_launch(w, out, rows, cols, ivg::MEGA_BN, S(stream));
}

long long ivgpt_mega_packed_elems_bn(int rows, int cols, int bn) {
  return bn > 0 ? (long long)((rows + bn - 1) / bn) * bn * cols : 0;
}

int ivgpt_mega_pack_weight_bn(const void* w, void* out, int rows, int cols, int bn, void* stream) {
  return ivg::mega_pack_weight_launch(w, out, rows, cols, bn, S(stream));
}

long long ivgpt_mega_packed_elems64(int rows, int cols) {
  return (long long)((rows + ivg::MEGA_WM - 1) / ivg::MEGA_WM) * ivg::MEGA_WM * cols;
}

int ivgpt_mega_pack_weight64(const void* w, void* out, int rows, int cols, int swiglu_pairs, void* stream) {
  return ivg::mega_pack_weight64_launch(w, out, rows, cols, swiglu_pairs, S(stream));
}

int ivgpt_mega_fill_layer(void* host_layer, const void* wqkv, const void* wo, const void* wgu, const void* wd,
                          const float* n1, const float* n2) {
  ivg::MegaLayer* L = reinterpret_cast<ivg::MegaLayer*>(host_layer);
  L->wqkv = (const __nv_bfloat16*)wqkv; L->wo = (const __nv_bfloat16*)wo;
  L->wgu = (const __nv_bfloat16*)wgu; L->wd = (const __nv_bfloat16*)wd;
  L->n1 = n1; L->n2 = n2;
  return 0;
}

int ivgpt_flash_attn(const void* q, const void* k, const void* vt, void* out, float* lse, int B, int heads, int Lq, int Lk,
                     long long q_bstride, long long k_bstride, long long vt_bstride, long long vt_ld, long long ldo, int causal,
                     float scale, void* stream) {
  return flash_attn_launch(q, k, vt, out, lse, B, heads, Lq, Lk, q_bstride, k_bstride, vt_bstride, vt_ld, ldo, causal, scale,
                           num_sms(), S(stream));
}

int ivgpt_set_deterministic(int on) { mma_issuers(); g_issuers = (on || !g_issuers_optin) ? 1 : 2; return 0; }
int ivgpt_set_gemm_mh2(int on) { g_mh2 = on ? 1 : 0; return 0; }
int ivgpt_set_mma_issuers(int n) { mma_issuers(); g_issuers_optin = n == 2 ? 1 : 0; g_issuers = n == 2 ? 2 : 1; return 0; }

int ivgpt_mega_fused_norm(void) { return ivg::mega_fused_norm(); }

int ivgpt_decode_mega(const ivgpt_mega_desc* d, void* stream) {
  IVG_CHECK(d != nullptr, "decode_mega: null descriptor");
  ivg::MegaParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.hidden = d->hidden; p.inter = d->inter; p.heads = d->heads; p.layers = d->layers; p.vocab = d->vocab;
  p.Lmax = d->Lmax; p.steps = d->steps; p.eps = d->eps; p.o_splits = d->o_splits; p.d_splits = d->d_splits;
  p.x = (float*)d->x; p.xn = (__nv_bfloat16*)d->xn; p.qkv = (__nv_bfloat16*)d->qkv; p.ao = (__nv_bfloat16*)d->ao;
  p.act = (__nv_bfloat16*)d->act; p.part = (float*)d->part; p.logits = (float*)d->logits; p.ldl = d->ldl;
  p.kcache = (__nv_bfloat16*)d->kcache; p.vcache = (__nv_bfloat16*)d->vcache; p.vrows = (__nv_bfloat16*)d->vrows;
  p.embed = d->embed; p.norm_f = d->norm_f; p.cos_tab = d->cos_tab; p.sin_tab = d->sin_tab;
  p.tokens = d->tokens; p.tok_stride = d->tok_stride; p.dpos = d->dpos;
  p.do_sample = d->do_sample; p.topk = d->topk; p.inv_temp = d->inv_temp; p.dseed = d->dseed;
  p.barrier = d->barrier; p.error = d->error;
  p.lw = reinterpret_cast<const ivg::MegaLayer*>(d->layers_dev);
  p.lm_head = reinterpret_cast<const __nv_bfloat16*>(d->lm_head_packed);
  p.prof = d->prof;
  p.attn_mode = d->attn_mode;
  p.attn_part = (float*)d->attn_part; p.attn_cnt = (unsigned int*)d->attn_cnt;
  p.tile_cnt = (unsigned int*)d->tile_cnt;
  p.slot_emb = d->slot_emb; p.slot0 = d->slot0; p.slot_period = d->slot_period; p.nslots = d->nslots;
  p.slot_token = d->slot_token;
  p.mma_m64 = 1;
  p.a_bulk = 1;
  IVG_CHECK(d->a_bulk == 1, "decode_mega: a_bulk must be 1 (xn / ao / act are swizzled activation images)");
  p.bn_wide = d->bn_wide > 0 ? d->bn_wide : ivg::MEGA_BN;
  p.bn_down = d->bn_down > 0 ? d->bn_down : ivg::MEGA_BN;
  p.gemm_mode = d->gemm_mode; p.qkv_splits = d->qkv_splits; p.qkvp = (float*)d->qkvp;
  p.a_rows = d->gemm_mode == 0 ? (d->B <= 64 ? 64 : 128) : d->a_rows;
  IVG_CHECK(p.slot_period >= 0 && (p.slot_period == 0 || p.nslots >= 1), "decode_mega: bad slot layout");
  IVG_CHECK(p.attn_mode == 1 || p.vrows != nullptr, "decode_mega: attn_mode 0 needs the row-major V cache (vrows)");
  IVG_CHECK(p.attn_mode == 1 || (p.attn_part != nullptr && p.attn_cnt != nullptr),
            "decode_mega: attn_mode 0 needs attn_part [SMs*4*72 floats] and attn_cnt [SMs uints, zeroed]");
  if (p.steps <= 0) return 0;
  const int rec = ivg::profile_begin(2, (double)p.steps, S(stream));      // bench.py roofline leg: work = decode steps
  const int rc = ivg::decode_mega_launch(p, num_sms(), S(stream));
  ivg::profile_end(rec, S(stream));
  return rc;
}

}  // extern "C"
