// extern "C" boundary: argument validation, TMA descriptor construction, launches.  See include/bevgen_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/bevgen_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

using namespace bevgen;

namespace bevgen { int g_pdl_enabled = 0; }

namespace {
thread_local char g_err[512] = "";
int g_sm_count = 0;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_mu;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int ensure_init() {
  if (g_encode != nullptr && g_sm_count > 0) return BEVGEN_OK;
  return bevgen_init(-1);
}

// bf16 tensor map (SWIZZLE_128B unless swizzle == false); dims/strides innermost first
int make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
              bool swizzle = true, bool swizzle64 = false) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  if (((uintptr_t)base & 15) != 0) return fail(BEVGEN_ERR_ARG, "tensor map base not 16B aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16 != 0) return fail(BEVGEN_ERR_ARG, "tensor map stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BEVGEN_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return BEVGEN_OK;
}
}  // namespace

extern "C" {

BEVGEN_API const char* bevgen_last_error(void) { return g_err; }
BEVGEN_API int bevgen_version(void) { return 100; }
BEVGEN_API int bevgen_sm_count(void) { return g_sm_count; }
BEVGEN_API int bevgen_set_pdl(int enabled) { const int old = bevgen::g_pdl_enabled; bevgen::g_pdl_enabled = enabled ? 1 : 0; return old; }

BEVGEN_API int bevgen_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return fail(BEVGEN_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail(BEVGEN_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(BEVGEN_ERR_ARCH, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
  g_sm_count = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
      qres != cudaDriverEntryPointSuccess)
    return fail(BEVGEN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  g_encode = (EncodeTiledFn)fn;
  return BEVGEN_OK;
}

BEVGEN_API int bevgen_gemm_tc(const bevgen_gemm_args* a, void* stream) {
  if (a == nullptr) return fail(BEVGEN_ERR_ARG, "null args");
  int rc = ensure_init();
  if (rc) return rc;
  const int bn = a->bn, npass = a->npass;
  if (!(bn == 16 || bn == 64 || bn == 128)) return fail(BEVGEN_ERR_ARG, "bn must be 16, 64 or 128 (got %d)", bn);
  if (!(npass >= 1 && npass <= 3)) return fail(BEVGEN_ERR_ARG, "npass must be 1, 2 or 3 (got %d)", npass);
  if (npass >= 2 && (a->a_lo == nullptr || a->b_lo == nullptr)) return fail(BEVGEN_ERR_ARG, "npass=%d needs lo / pair planes", npass);
  if (npass == 2 && (bn != 128 || (a->flags & BEVGEN_GF_B_MN) || !(a->lo_scale > 0.f)))
    return fail(BEVGEN_ERR_ARG, "npass=2 (f16f8) needs bn = 128, a K-major B operand and lo_scale > 0");
  if ((a->flags & BEVGEN_GF_OUT_F16F8) && (!a->out_hi || !a->out_lo || (a->ldc & 63) || (a->n_cols & 31) || (a->flags & BEVGEN_GF_OUT_NCHW)))
    return fail(BEVGEN_ERR_ARG, "f16f8 output planes need out_hi + out_lo, ldc %% 64 == 0 and n_cols %% 32 == 0");
  if (a->ntaps < 1 || a->ntaps > BEVGEN_MAX_TAPS) return fail(BEVGEN_ERR_ARG, "ntaps %d out of range", a->ntaps);
  if (a->k <= 0 || a->k % 64 != 0) return fail(BEVGEN_ERR_ARG, "k (%d) must be a positive multiple of 64", a->k);
  if (a->tile_w * a->tile_h != 128 || a->tile_w > 256 || a->tile_h > 256) return fail(BEVGEN_ERR_ARG, "tile %dx%d != 128 pixels", a->tile_w, a->tile_h);
  if (a->a_c % 8 != 0 || a->b_cols % 8 != 0) return fail(BEVGEN_ERR_ARG, "channel counts must be multiples of 8");
  if (a->z_inner < 1 || a->z_outer < 1 || a->out_w < 1 || a->out_h < 1 || a->n_cols < 1) return fail(BEVGEN_ERR_ARG, "empty problem");
  if ((a->flags & BEVGEN_GF_B_MN) && bn < 64) return fail(BEVGEN_ERR_ARG, "MN-major B needs bn >= 64");
  if (a->out_f32 == nullptr && a->out_hi == nullptr) return fail(BEVGEN_ERR_ARG, "no output buffer");
  if ((a->flags & BEVGEN_GF_OUT_NCHW) && a->out_hi != nullptr) return fail(BEVGEN_ERR_ARG, "NCHW output is fp32 only");
  if ((a->flags & BEVGEN_GF_OUT_T) && (a->out_f32 == nullptr || a->out_hi != nullptr || a->bias || a->residual || (a->flags & BEVGEN_GF_GELU)))
    return fail(BEVGEN_ERR_ARG, "transposed output is a raw fp32 partial (no bias/activation/residual/split planes)");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  const void* aplanes[2] = {a->a_hi, a->a_lo};
  const void* bplanes[2] = {a->b_hi, a->b_lo};
  for (int o = 0; o < (npass >= 2 ? 2 : 1); ++o) {
    uint64_t ad[4] = {(uint64_t)a->a_c, (uint64_t)a->a_w, (uint64_t)a->a_h, (uint64_t)a->a_n};
    uint64_t as[3] = {(uint64_t)a->a_c * 2, (uint64_t)a->a_c * a->a_w * 2, (uint64_t)a->a_c * a->a_w * a->a_h * 2};
    uint32_t ab[4] = {64, (uint32_t)a->tile_w, (uint32_t)a->tile_h, 1};
    rc = make_tmap(&p.tmA[o], aplanes[o], 4, ad, as, ab);
    if (rc) return rc;
    uint64_t bd[2] = {(uint64_t)a->b_cols, (uint64_t)a->b_rows};
    uint64_t bs[1] = {(uint64_t)a->b_cols * 2};
    uint32_t bb[2] = {64, (uint32_t)((a->flags & BEVGEN_GF_B_MN) ? 64 : bn)};
    rc = make_tmap(&p.tmB[o], bplanes[o], 2, bd, bs, bb);
    if (rc) return rc;
  }
  p.ntaps = a->ntaps;
  for (int t = 0; t < a->ntaps; ++t) { p.tap_dx[t] = a->tap_dx[t]; p.tap_dy[t] = a->tap_dy[t]; p.tap_dn[t] = a->tap_dn[t]; }
  p.a_n_mul = a->a_n_mul; p.a_n_zstride = a->a_n_zstride;
  p.kchunks = a->k / 64;
  p.a_c_off = a->a_c_off; p.a_c_zstride = a->a_c_zstride;
  p.b_k_off = a->b_k_off; p.b_k_zstride = a->b_k_zstride;
  p.b_row_zstride = a->b_row_zstride; p.b_row_tapstride = a->b_row_tapstride;
  p.z_inner = a->z_inner; p.z_outer = a->z_outer;
  p.tile_w = a->tile_w; p.tile_h = a->tile_h;
  p.tiles_w = (a->out_w + a->tile_w - 1) / a->tile_w;
  p.tiles_h = (a->out_h + a->tile_h - 1) / a->tile_h;
  p.out_w = a->out_w; p.out_h = a->out_h; p.n_cols = a->n_cols;
  p.out_zo_stride = a->out_zo_stride; p.out_zi_stride = a->out_zi_stride; p.ldc = a->ldc;
  p.bias = a->bias; p.residual = a->residual; p.out_f32 = a->out_f32;
  p.out_hi = (uint16_t*)a->out_hi; p.out_lo = (uint16_t*)a->out_lo;
  p.flags = a->flags; p.causal_ncond = a->causal_ncond; p.lo_scale = a->lo_scale;
  p.fin_mode = a->fin_mode;
  if (a->fin_mode != 0) {
    if (!(a->flags & BEVGEN_GF_OUT_T) || a->z_outer != 1 || !a->fin_counters || !a->fin_hi || a->fin_rows < 1 || a->fin_rows > a->n_cols)
      return fail(BEVGEN_ERR_ARG, "fused finalize needs a transposed split-K launch, counters and output planes");
    if (a->fin_mode == 2 && (!a->fin_resid || !a->fin_x || !a->fin_gamma || !a->fin_beta || (a->out_w * a->out_h) % 128 != 0 || a->out_w * a->out_h > 1024))
      return fail(BEVGEN_ERR_ARG, "residual+LayerNorm finalize: missing buffers or feature count not a multiple of 128 <= 1024");
    p.fin_gelu = a->fin_gelu; p.fin_rows = a->fin_rows; p.fin_bias = a->fin_bias; p.fin_resid = a->fin_resid; p.fin_x = a->fin_x;
    p.fin_y = a->fin_y; p.fin_hi = (uint16_t*)a->fin_hi; p.fin_lo = (uint16_t*)a->fin_lo; p.fin_gamma = a->fin_gamma; p.fin_beta = a->fin_beta;
    p.fin_eps = a->fin_eps; p.fin_counters = a->fin_counters;
  }
  rc = gemm_tc_dispatch(p, bn, npass, g_sm_count, (cudaStream_t)stream);
  if (rc) return fail(rc, "gemm_tc launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return BEVGEN_OK;
}

#define CHECK_LAUNCH_DEFINED 1
#define CHECK_LAUNCH(expr, name)                                                                         \
  do {                                                                                                   \
    int rc_ = (expr);                                                                                    \
    if (rc_) return fail(rc_, "%s failed (%d): %s", name, rc_, cudaGetErrorString(cudaGetLastError())); \
    return BEVGEN_OK;                                                                                    \
  } while (0)

BEVGEN_API int bevgen_conv3x3_halo(const void* a_hi, const void* a_lo, int n, int h, int w, int cin, const void* w_hi, const void* w_lo, int w_rows,
                                   int cout, const float* bias, const float* residual, float* out, double* gn_sums, int npass, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!a_hi || !w_hi || !bias || !out || (npass == 3 && (!a_lo || !w_lo)) || !(npass == 1 || npass == 3)) return fail(BEVGEN_ERR_ARG, "conv3x3_halo: bad args");
  if (cin % 64 != 0 || cout % 32 != 0 || n < 1 || h < 1 || w < 1) return fail(BEVGEN_ERR_ARG, "conv3x3_halo: cin %% 64 / cout %% 32 required (cin=%d cout=%d)", cin, cout);
  if (w_rows < 8 * cout + ((cout + 127) / 128) * 128) return fail(BEVGEN_ERR_ARG, "conv3x3_halo: weight rows must be padded to 8*cout + ceil128(cout)");
  ConvHaloParams p;
  memset(&p, 0, sizeof(p));
  const void* ap[2] = {a_hi, a_lo};
  const void* wp[2] = {w_hi, w_lo};
  for (int o = 0; o < (npass == 3 ? 2 : 1); ++o) {
    uint64_t ad[5] = {8, (uint64_t)w, (uint64_t)h, (uint64_t)cin / 8, (uint64_t)n};
    uint64_t as[4] = {(uint64_t)cin * 2, (uint64_t)cin * w * 2, 16, (uint64_t)cin * w * h * 2};
    uint32_t ab[5] = {8, 10, 18, 8, 1};
    rc = make_tmap(&p.tmA[o], ap[o], 5, ad, as, ab, false);
    if (rc) return rc;
    uint64_t wd[2] = {(uint64_t)cin, (uint64_t)w_rows};
    uint64_t wst[1] = {(uint64_t)cin * 2};
    uint32_t wb[2] = {64, 128};
    rc = make_tmap(&p.tmW[o], wp[o], 2, wd, wst, wb, true);
    if (rc) return rc;
  }
  p.N = n; p.H = h; p.W = w; p.Cin = cin; p.Cout = cout;
  p.bias = bias; p.residual = residual; p.out = out; p.gn_sums = gn_sums;
  if (gn_sums != nullptr && cudaMemsetAsync(gn_sums, 0, (size_t)n * 64 * sizeof(double), (cudaStream_t)stream) != cudaSuccess)
    return fail(BEVGEN_ERR_CUDA, "conv3x3_halo: memset failed");
  CHECK_LAUNCH(launch_conv_halo(p, npass, g_sm_count, (cudaStream_t)stream), "conv3x3_halo");
}

static int conv3x3_fused_impl(const float* x, int n, int h, int w, int cin, const float* affine, int swish, int up2, const void* w_hi,
                              const void* w_lo, int w_rows, int cout, const float* bias, const float* residual, float* out, double* gn_sums,
                              int npass, float lo_scale, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  const bool two_cta = (npass & 0x100) != 0;      // bit 8 of npass selects the cta_group::2 (cluster of two CTAs) kernel
  const bool block16 = (npass & 0x200) != 0;      // bit 9: weight-stationary 16x16-block kernel (conv_fused3.cu), 32-channel weight boxes
  npass &= 0xff;
  if (!x || !w_hi || !bias || !out || (npass >= 2 && !w_lo) || !(npass >= 1 && npass <= 3)) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: bad args");
  if (npass == 2 && ((!two_cta && !block16) || !(lo_scale > 0.f))) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: the f16+f8 mode exists for the 2-CTA kernels only and needs lo_scale > 0");
  if (block16 && cin % 32 != 0) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: the block kernel needs cin %% 32 == 0");
  if (!block16 && cin % 64 != 0) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: cin %% 64 required (cin=%d)", cin);
  if (cin % 32 != 0 || cout % 32 != 0 || n < 1 || h < 1 || w < 1) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: cin %% 32 / cout %% 32 required (cin=%d cout=%d)", cin, cout);
  if (w_rows < 8 * cout + ((cout + 127) / 128) * 128) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: weight rows must be padded to 8*cout + ceil128(cout)");
  if (((uintptr_t)x & 15) != 0 || (affine && ((uintptr_t)affine & 15) != 0)) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: x / affine must be 16-byte aligned");
  ConvFusedParams p;
  memset(&p, 0, sizeof(p));
  const void* wp[2] = {w_hi, w_lo};
  for (int o = 0; o < (npass >= 2 ? 2 : 1); ++o) {      // npass == 2: plane 1 is the packed e4m3 pair, also 2*cin bytes per row
    uint64_t wd[2] = {(uint64_t)cin, (uint64_t)w_rows};
    uint64_t wst[1] = {(uint64_t)cin * 2};
    uint32_t wb[2] = {(uint32_t)(block16 ? 32 : 64), (uint32_t)((two_cta || block16) ? 64 : 128)};
    rc = make_tmap(&p.tmW[o], wp[o], 2, wd, wst, wb, true, block16);
    if (rc) return rc;
  }
  p.x = x; p.affine = affine; p.swish = swish; p.up2 = up2;
  p.N = n; p.H = h; p.W = w; p.Cin = cin; p.Cout = cout;
  p.bias = bias; p.residual = residual; p.out = out; p.gn_sums = gn_sums; p.lo_scale = lo_scale;
  { const char* e = getenv("BEVGEN_CONV_DBG"); p.dbg = e ? atoi(e) : 0; }
  if (gn_sums != nullptr && cudaMemsetAsync(gn_sums, 0, (size_t)n * 64 * sizeof(double), (cudaStream_t)stream) != cudaSuccess)
    return fail(BEVGEN_ERR_CUDA, "conv3x3_fused: memset failed");
  if (block16) {
    if (npass != 2) p.lo_scale = 1.0f;            // the block kernel applies lo_scale to the whole accumulator
    CHECK_LAUNCH(launch_conv_fused3(p, npass, g_sm_count, (cudaStream_t)stream), "conv3x3_fused (16x16 block)");
  }
  if (two_cta) CHECK_LAUNCH(launch_conv_fused2(p, npass, g_sm_count, (cudaStream_t)stream), "conv3x3_fused (2-CTA)");
  CHECK_LAUNCH(launch_conv_fused(p, npass, g_sm_count, (cudaStream_t)stream), "conv3x3_fused");
}

BEVGEN_API int bevgen_conv3x3_fused(const float* x, int n, int h, int w, int cin, const float* affine, int swish, int up2, const void* w_hi,
                                    const void* w_lo, int w_rows, int cout, const float* bias, const float* residual, float* out, double* gn_sums,
                                    int npass, void* stream) {
  if ((npass & 0xff) == 2) return fail(BEVGEN_ERR_ARG, "conv3x3_fused: npass 2 is bevgen_conv3x3_fused_f16f8");
  return conv3x3_fused_impl(x, n, h, w, cin, affine, swish, up2, w_hi, w_lo, w_rows, cout, bias, residual, out, gn_sums, npass, 0.f, stream);
}

BEVGEN_API int bevgen_conv3x3_fused_f16f8(const float* x, int n, int h, int w, int cin, const float* affine, int swish, int up2, const void* w_f16,
                                          const void* w_f8pair, int w_rows, int cout, float lo_scale, const float* bias, const float* residual,
                                          float* out, double* gn_sums, int block16, void* stream) {
  return conv3x3_fused_impl(x, n, h, w, cin, affine, swish, up2, w_f16, w_f8pair, w_rows, cout, bias, residual, out, gn_sums,
                            2 | (block16 ? 0x200 : 0x100), lo_scale, stream);
}

BEVGEN_API int bevgen_groupnorm_affine(const double* sums, const float* gamma, const float* beta, int n, int pixels, int c, float eps, float* affine,
                                       void* stream) {
  if (!sums || !gamma || !beta || !affine || n < 1 || c % 32 != 0) return fail(BEVGEN_ERR_ARG, "groupnorm_affine: bad args");
  CHECK_LAUNCH(launch_gn_affine(sums, gamma, beta, affine, n, pixels, c, eps, (cudaStream_t)stream), "groupnorm_affine");
}

BEVGEN_API int bevgen_groupnorm_finalize(const double* sums, int n, int pixels, int c, float eps, float* mean_rstd, void* stream) {
  if (!sums || !mean_rstd || n < 1 || c % 32 != 0) return fail(BEVGEN_ERR_ARG, "groupnorm_finalize: bad args");
  CHECK_LAUNCH(launch_gn_finalize(sums, mean_rstd, n, pixels, c, eps, (cudaStream_t)stream), "groupnorm_finalize");
}

BEVGEN_API int bevgen_groupnorm_stats(const float* x, int n, int pixels, int c, float eps, double* ws_sums, float* mean_rstd, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x || !ws_sums || !mean_rstd || n < 1 || pixels < 1) return fail(BEVGEN_ERR_ARG, "groupnorm_stats: bad args");
  CHECK_LAUNCH(launch_gn_stats(x, ws_sums, mean_rstd, n, pixels, c, eps, (cudaStream_t)stream), "groupnorm_stats");
}

BEVGEN_API int bevgen_prep_operand(const float* x, int n, int h, int w, int c, const float* mean_rstd, const float* gamma, const float* beta,
                        int swish, int mode, void* out_hi, void* out_lo, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x || !out_hi || n < 1 || h < 1 || w < 1) return fail(BEVGEN_ERR_ARG, "prep_operand: bad args");
  if (mean_rstd && (!gamma || !beta)) return fail(BEVGEN_ERR_ARG, "prep_operand: affine params missing");
  if (mode < 0 || mode > 2) return fail(BEVGEN_ERR_ARG, "prep_operand: bad mode");
  PrepParams p{x, mean_rstd, gamma, beta, (uint16_t*)out_hi, (uint16_t*)out_lo, n, h, w, c, mode, swish};
  CHECK_LAUNCH(launch_prep(p, g_sm_count, (cudaStream_t)stream), "prep_operand");
}

BEVGEN_API int bevgen_im2col3x3(const float* x_nchw, int n, int cin, int h, int w, void* out_hi, void* out_lo, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x_nchw || !out_hi) return fail(BEVGEN_ERR_ARG, "im2col3x3: bad args");
  CHECK_LAUNCH(launch_im2col3x3(x_nchw, (uint16_t*)out_hi, (uint16_t*)out_lo, n, cin, h, w, g_sm_count, (cudaStream_t)stream), "im2col3x3");
}

BEVGEN_API int bevgen_transpose_f32(const float* src, float* dst, int n, int r, int c, void* stream) {
  if (!src || !dst || n < 1 || r < 1 || c < 1) return fail(BEVGEN_ERR_ARG, "transpose: bad args");
  CHECK_LAUNCH(launch_transpose(src, dst, n, r, c, (cudaStream_t)stream), "transpose_f32");
}

BEVGEN_API int bevgen_softmax_rows(const float* s, long long rows, int cols, float scale, void* out_hi, void* out_lo, int out_ld, void* stream) {
  if (!s || !out_hi || rows < 1 || out_ld < cols) return fail(BEVGEN_ERR_ARG, "softmax_rows: bad args");
  CHECK_LAUNCH(launch_softmax_rows(s, (uint16_t*)out_hi, (uint16_t*)out_lo, rows, cols, out_ld, scale, (cudaStream_t)stream), "softmax_rows");
}

BEVGEN_API int bevgen_row_sqnorm(const float* x, int rows, int dim, float* out, void* stream) {
  if (!x || !out || rows < 1 || dim < 1) return fail(BEVGEN_ERR_ARG, "row_sqnorm: bad args");
  CHECK_LAUNCH(launch_row_sqnorm(x, out, rows, dim, (cudaStream_t)stream), "row_sqnorm");
}

BEVGEN_API int bevgen_vq_nearest(const float* z, const float* codebook, const float* code_sqnorm, int rows, int n_codes, int dim, float* ws_zz,
                      long long* idx, float* zq, void* stream) {
  if (!z || !codebook || !code_sqnorm || !ws_zz || !idx || rows < 1 || n_codes < 1) return fail(BEVGEN_ERR_ARG, "vq_nearest: bad args");
  int rc = launch_row_sqnorm(z, ws_zz, rows, dim, (cudaStream_t)stream);
  if (rc) return fail(rc, "vq_nearest: row_sqnorm failed");
  CHECK_LAUNCH(launch_vq_nearest(z, codebook, ws_zz, code_sqnorm, idx, zq, rows, n_codes, dim, (cudaStream_t)stream), "vq_nearest");
}

BEVGEN_API int bevgen_codebook_gather(const float* codebook, const long long* idx, long long rows, int dim, int n_codes, float* out, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!codebook || !idx || !out || rows < 1) return fail(BEVGEN_ERR_ARG, "codebook_gather: bad args");
  CHECK_LAUNCH(launch_gather_rows(codebook, idx, out, rows, dim, n_codes, g_sm_count, (cudaStream_t)stream), "codebook_gather");
}

BEVGEN_API int bevgen_denormalize(const float* x, float* out, int n, int c, int pixels, const float* mean3, const float* std3, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x || !out || !mean3 || !std3) return fail(BEVGEN_ERR_ARG, "denormalize: bad args");
  CHECK_LAUNCH(launch_denorm(x, out, n, c, pixels, mean3, std3, g_sm_count, (cudaStream_t)stream), "denormalize");
}

BEVGEN_API int bevgen_conv_in3(const float* x_nchw, const float* weight_oihw, const float* bias, float* out_nhwc, double* gn_sums, int n, int h, int w,
                               int cout, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x_nchw || !weight_oihw || !out_nhwc) return fail(BEVGEN_ERR_ARG, "conv_in3: bad args");
  if (!(cout == 64 || cout == 128)) return fail(BEVGEN_ERR_ARG, "conv_in3: cout must be 64 or 128 (got %d)", cout);
  if (gn_sums != nullptr && cudaMemsetAsync(gn_sums, 0, (size_t)n * 64 * sizeof(double), (cudaStream_t)stream) != cudaSuccess)
    return fail(BEVGEN_ERR_CUDA, "conv_in3: memset failed");
  CHECK_LAUNCH(launch_conv_in3(x_nchw, weight_oihw, bias, out_nhwc, gn_sums, n, h, w, cout, g_sm_count, (cudaStream_t)stream), "conv_in3");
}

BEVGEN_API int bevgen_conv_out3(const float* x_nhwc, int n, int h, int w, int c, const float* affine, int swish, const float* weight_oihw,
                                const float* bias, float* out_nchw, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x_nhwc || !weight_oihw || !out_nchw) return fail(BEVGEN_ERR_ARG, "conv_out3: bad args");
  if (!(c == 64 || c == 128)) return fail(BEVGEN_ERR_ARG, "conv_out3: c must be 64 or 128 (got %d)", c);
  if (((uintptr_t)x_nhwc & 15) != 0) return fail(BEVGEN_ERR_ARG, "conv_out3: x must be 16-byte aligned");
  CHECK_LAUNCH(launch_conv_out3(x_nhwc, affine, swish, weight_oihw, bias, out_nchw, n, h, w, c, g_sm_count, (cudaStream_t)stream), "conv_out3");
}

BEVGEN_API int bevgen_absmax(const float* x, long long n, float* out, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x || !out || n < 1) return fail(BEVGEN_ERR_ARG, "absmax: bad args");
  CHECK_LAUNCH(launch_absmax(x, n, out, g_sm_count, (cudaStream_t)stream), "absmax");
}

BEVGEN_API int bevgen_pack_split_bf16(const float* x, long long n, void* hi, void* lo, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x || !hi || n < 1) return fail(BEVGEN_ERR_ARG, "pack_split_bf16: bad args");
  CHECK_LAUNCH(launch_split_bf16(x, n, hi, lo, g_sm_count, (cudaStream_t)stream), "pack_split_bf16");
}

BEVGEN_API int bevgen_pack_f16f8(const float* w, long long rows, int cin, int chunk, float s, float w16_mul, void* w16, void* pair, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!w || !w16 || !pair) return fail(BEVGEN_ERR_ARG, "pack_f16f8: null pointer");
  CHECK_LAUNCH(launch_pack_f16f8(w, rows, cin, chunk, s, w16_mul, w16, pair, g_sm_count, (cudaStream_t)stream), "pack_f16f8");
}

BEVGEN_API int bevgen_to_uint8_hwc(const float* x_nchw, void* out_nhwc_u8, int n, int c, int pixels, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!x_nchw || !out_nhwc_u8) return fail(BEVGEN_ERR_ARG, "to_uint8_hwc: bad args");
  CHECK_LAUNCH(launch_to_uint8_hwc(x_nchw, (uint8_t*)out_nhwc_u8, n, c, pixels, g_sm_count, (cudaStream_t)stream), "to_uint8_hwc");
}

BEVGEN_API int bevgen_layernorm(const float* x, long long rows, int d, long long x_row_stride, const float* gamma, const float* beta, float eps,
                                float* y, void* out_hi, void* out_lo, void* stream) {
  if (!x || !gamma || !beta || (!y && !out_hi)) return fail(BEVGEN_ERR_ARG, "layernorm: bad args");
  if (x_row_stride % 4 != 0) return fail(BEVGEN_ERR_ARG, "layernorm: row stride must be a multiple of 4");
  CHECK_LAUNCH(launch_layernorm(x, gamma, beta, y, (uint16_t*)out_hi, (uint16_t*)out_lo, rows, d, x_row_stride, eps, 0, (cudaStream_t)stream), "layernorm");
}

BEVGEN_API int bevgen_layernorm_f16f8(const float* x, long long rows, int d, long long x_row_stride, const float* gamma, const float* beta, float eps,
                                      float* y, void* out_f16, void* out_f8pair, int scaled, void* stream) {
  if (!x || !gamma || !beta || !out_f16 || !out_f8pair) return fail(BEVGEN_ERR_ARG, "layernorm_f16f8: bad args");
  if (x_row_stride % 4 != 0) return fail(BEVGEN_ERR_ARG, "layernorm: row stride must be a multiple of 4");
  CHECK_LAUNCH(launch_layernorm(x, gamma, beta, y, (uint16_t*)out_f16, (uint16_t*)out_f8pair, rows, d, x_row_stride, eps, scaled ? 2 : 1,
                                (cudaStream_t)stream), "layernorm_f16f8");
}

BEVGEN_API int bevgen_linear_f16f8(const void* a16, const void* apair, const void* w16, const void* wpair, long long M, int N, int K, float out_scale,
                                   const float* bias, int gelu, const float* residual, float* out_f32, void* out_hi, void* out_lo, void* out_f16,
                                   void* out_pair, void* stream) {
  if (ensure_init() != BEVGEN_OK) return BEVGEN_ERR_DRIVER;
  if (!a16 || !apair || !w16 || !wpair || M < 1 || M > 0x7fffffffLL) return fail(BEVGEN_ERR_ARG, "linear_f16f8: bad args");
  if (N < 32 || N % 32 != 0 || K < 64 || K % 64 != 0) return fail(BEVGEN_ERR_ARG, "linear_f16f8: need N %% 32 == 0 and K %% 64 == 0 (N=%d K=%d)", N, K);
  if ((out_hi == nullptr) != (out_lo == nullptr) || (out_f16 == nullptr) != (out_pair == nullptr) || (!out_f32 && !out_hi && !out_f16))
    return fail(BEVGEN_ERR_ARG, "linear_f16f8: outputs must be fp32 and/or a complete pair of planes");
  bevgen::GemmPairParams p{};
  const void* ap[2] = {a16, apair};
  const void* wp[2] = {w16, wpair};
  const uint64_t ad[2] = {(uint64_t)K, (uint64_t)M}, wd[2] = {(uint64_t)K, (uint64_t)N}, st[1] = {(uint64_t)K * 2};
  const uint32_t box[2] = {64, 128};
  int rc;
  for (int o = 0; o < 2; ++o) {
    rc = make_tmap(&p.tmA[o], ap[o], 2, ad, st, box);
    if (rc != BEVGEN_OK) return rc;
    rc = make_tmap(&p.tmW[o], wp[o], 2, wd, st, box);
    if (rc != BEVGEN_OK) return rc;
  }
  p.M = (int)M; p.N = N; p.K = K; p.out_scale = out_scale; p.bias = bias; p.gelu = gelu; p.residual = residual; p.out_f32 = out_f32;
  p.out_hi = (uint16_t*)out_hi; p.out_lo = (uint16_t*)out_lo; p.out_f16 = (uint16_t*)out_f16; p.out_pair = out_pair;
  CHECK_LAUNCH(bevgen::launch_gemm_pair_f16f8(p, g_sm_count, (cudaStream_t)stream), "linear_f16f8");
}

BEVGEN_API int bevgen_embed_assemble(const bevgen_embed_args* a, void* stream) {
  if (!a || !a->cam_idx || !a->bev_idx || !a->x_tok_emb || !a->cond_tok_emb || !a->x_pos_emb || !a->cond_static || !a->forward_shuffle_idx || !a->out)
    return fail(BEVGEN_ERR_ARG, "embed_assemble: null argument");
  if (a->img_embed_w && (!a->cam_embed_w || !a->intrinsics_inv || !a->extrinsics_inv || !a->pixel)) return fail(BEVGEN_ERR_ARG, "embed_assemble: ray embedding inputs missing");
  if (!a->step_ptr && (a->row0 < 0 || a->row0 + a->nrows > a->L)) return fail(BEVGEN_ERR_ARG, "embed_assemble: row range outside the sequence");
  EmbedParams p{a->cam_idx, a->bev_idx, a->intrinsics_inv, a->extrinsics_inv, a->x_tok_emb, a->cond_tok_emb, a->x_pos_emb, a->cond_static,
                a->img_embed_w, a->cam_embed_w, a->forward_shuffle_idx, a->pixel, a->out, a->step_ptr, a->B, a->ncam, a->hw, a->nc, a->n_img, a->L, a->d,
                a->vocab, a->pad_last, a->bev_embed, a->row0, a->nrows};
  CHECK_LAUNCH(launch_embed(p, (cudaStream_t)stream), "embed_assemble");
}

BEVGEN_API int bevgen_attn_softmax(const float* s, const float* bias, const unsigned char* mask, long long zrows, int L, int Lk, float scale,
                                   void* out_hi, void* out_lo, const unsigned char* layout, int heads, int block, int layout_ld, void* stream) {
  if (!s || !mask || !out_hi) return fail(BEVGEN_ERR_ARG, "attn_softmax: bad args");
  CHECK_LAUNCH(launch_attn_softmax(s, bias, mask, (uint16_t*)out_hi, (uint16_t*)out_lo, zrows, L, Lk, scale, layout, heads, block, layout_ld,
                                   (cudaStream_t)stream), "attn_softmax");
}

BEVGEN_API int bevgen_attn_fused_fwd(const void* qkv_hi, const void* qkv_lo, int batch, int seq_len, int heads, int d, int n_cond,
                                     const void* bias_f16, const float* y, float* x1, float scale, int npass, const unsigned long long* layout64,
                                     void* out_hi, void* out_lo, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!qkv_hi || (!x1 && !out_hi) || (npass == 3 && !qkv_lo) || !(npass == 1 || npass == 3)) return fail(BEVGEN_ERR_ARG, "attn_fused_fwd: bad args");
  if (seq_len % 128 != 0 || n_cond % 128 != 0 || n_cond < 128 || n_cond > seq_len || d != heads * 64)
    return fail(BEVGEN_ERR_ARG, "attn_fused_fwd: needs seq_len, n_cond multiples of 128 and d_head = 64 (got L=%d nc=%d d=%d H=%d)", seq_len, n_cond, d, heads);
  if (seq_len > 4096) return fail(BEVGEN_ERR_ARG, "attn_fused_fwd: seq_len %d > 4096", seq_len);
  CUtensorMap tm[2];
  const void* planes[2] = {qkv_hi, qkv_lo};
  for (int o = 0; o < (npass == 3 ? 2 : 1); ++o) {
    uint64_t dims[2] = {(uint64_t)3 * d, (uint64_t)batch * seq_len};
    uint64_t strides[1] = {(uint64_t)3 * d * 2};
    uint32_t box[2] = {64, 128};
    rc = make_tmap(&tm[o], planes[o], 2, dims, strides, box);
    if (rc) return rc;
  }
  CHECK_LAUNCH(launch_attn_fused(&tm[0], npass == 3 ? &tm[1] : nullptr, bias_f16, y, x1, batch, heads, seq_len, n_cond, d, scale, npass,
                                 layout64, (uint16_t*)out_hi, (uint16_t*)out_lo, (cudaStream_t)stream), "attn_fused_fwd");
}

BEVGEN_API int bevgen_ray_embed_add(float* h_nhwc, const float* intrinsics_inv, const float* extrinsics_inv, const float* pixel, const float* img_embed_w,
                                    const float* cam_embed_w, int n_images, int hw, int d, void* stream) {
  if (!h_nhwc || !intrinsics_inv || !extrinsics_inv || !pixel || !img_embed_w || !cam_embed_w) return fail(BEVGEN_ERR_ARG, "ray_embed_add: null argument");
  CHECK_LAUNCH(launch_ray_embed_add(h_nhwc, intrinsics_inv, extrinsics_inv, pixel, img_embed_w, cam_embed_w, n_images, hw, d, (cudaStream_t)stream),
               "ray_embed_add");
}

/* ---------------------------------------------------------------- MaskGit variant */
BEVGEN_API int bevgen_mg_head_planes(const float* src, long long src_ld, int src_col0, int n_src, int src_batch_rows, const float* null_vec,
                                     const float* scale, void* out_hi, void* out_lo, int batch, int dst_rows, int dst_batch_rows, long long dst_ld,
                                     int dst_col0, int has_null, int heads, void* stream) {
  if (!src || !out_hi) return fail(BEVGEN_ERR_ARG, "mg_head_planes: bad args");
  CHECK_LAUNCH(launch_mg_head_planes(src, src_ld, src_col0, n_src, src_batch_rows, null_vec, scale, (uint16_t*)out_hi, (uint16_t*)out_lo, batch,
                                     dst_rows, dst_batch_rows, dst_ld, dst_col0, has_null, heads, (cudaStream_t)stream), "mg_head_planes");
}

BEVGEN_API int bevgen_mg_sample(const float* logits, const float* uniform, long long* ids, float* scores, long long rows, int vocab, int top_k,
                                float inv_temperature, long long mask_id, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!logits || !uniform || !ids) return fail(BEVGEN_ERR_ARG, "mg_sample: null pointer");
  CHECK_LAUNCH(launch_mg_sample(logits, uniform, ids, scores, rows, vocab, top_k, inv_temperature, mask_id, (cudaStream_t)stream), "mg_sample");
}

BEVGEN_API int bevgen_mg_remask(const float* scores, const float* uniform, float noise_scale, long long* ids, const long long* init_ids, long long rows,
                                int hw, int n_mask, long long mask_id, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!scores || !ids) return fail(BEVGEN_ERR_ARG, "mg_remask: null pointer");
  CHECK_LAUNCH(launch_mg_remask(scores, uniform, noise_scale, ids, init_ids, rows, hw, n_mask, mask_id, (cudaStream_t)stream), "mg_remask");
}

BEVGEN_API int bevgen_mg_geglu_ln(const float* h, long long h_ld, const float* gamma, void* out_hi, void* out_lo, long long rows, int f, int f_pad,
                                  float eps, int f16f8, void* stream) {
  if (!h || !gamma || !out_hi) return fail(BEVGEN_ERR_ARG, "mg_geglu_ln: bad args");
  CHECK_LAUNCH(launch_mg_geglu_ln(h, h_ld, gamma, (uint16_t*)out_hi, (uint16_t*)out_lo, rows, f, f_pad, eps, f16f8, (cudaStream_t)stream), "mg_geglu_ln");
}

/* ---------------------------------------------------------------- KV-cache decode */
BEVGEN_API int bevgen_dec_reduce_ln(const float* partials, int ks, long long zstride, const float* bias, const float* residual,
                                    long long residual_row_stride, const float* gamma, const float* beta, float eps, float* x_out, float* y,
                                    void* out_hi, void* out_lo, int rows, int d, void* stream) {
  if (!gamma || !beta || (ks > 0 && !partials) || (ks == 0 && !residual)) return fail(BEVGEN_ERR_ARG, "dec_reduce_ln: bad args");
  CHECK_LAUNCH(launch_dec_reduce_ln(partials, ks, zstride, bias, residual, residual_row_stride, gamma, beta, eps, x_out, y, (uint16_t*)out_hi,
                                    (uint16_t*)out_lo, rows, d, (cudaStream_t)stream), "dec_reduce_ln");
}

BEVGEN_API int bevgen_dec_reduce_act(const float* partials, int ks, long long zstride, const float* bias, int gelu, void* out_hi, void* out_lo,
                                     int rows, int n, void* stream) {
  if (!partials || !bias || !out_hi || ks < 1) return fail(BEVGEN_ERR_ARG, "dec_reduce_act: bad args");
  CHECK_LAUNCH(launch_dec_reduce_act(partials, ks, zstride, bias, (uint16_t*)out_hi, (uint16_t*)out_lo, rows, n, gelu, (cudaStream_t)stream),
               "dec_reduce_act");
}

BEVGEN_API int bevgen_kv_store(const void* qkv_hi, const void* qkv_lo, void* k_cache, void* v_cache, int kv_bf16, int batch, int lp, int nrows,
                               int heads, int d, int lmax, void* stream) {
  if (!qkv_hi || !k_cache || !v_cache || nrows > lp || nrows > lmax) return fail(BEVGEN_ERR_ARG, "kv_store: bad args");
  CHECK_LAUNCH(launch_kv_store((const uint16_t*)qkv_hi, (const uint16_t*)qkv_lo, k_cache, v_cache, kv_bf16, batch, lp, nrows, heads, d, lmax,
                               (cudaStream_t)stream), "kv_store");
}

BEVGEN_API int bevgen_dec_attention(const float* qkv_partials, int ks, long long zstride, const float* qkv_bias, const float* y,
                                    const float* camera_bias, int bias_ld, void* k_cache, void* v_cache, int kv_bf16, float* x1,
                                    const int* step_ptr, float* workspace, unsigned int* counters, int batch, int n_cond, int heads, int d,
                                    int lmax, float scale, unsigned int* row_counters, const float* ln_gamma, const float* ln_beta, float ln_eps,
                                    void* ln_hi, void* ln_lo, const unsigned char* layout, int layout_block, int layout_ld, void* stream) {
  int rc0 = ensure_init();
  if (rc0) return rc0;
  if (!qkv_partials || !qkv_bias || !y || !k_cache || !v_cache || !x1 || !step_ptr || !workspace || !counters || ks < 1)
    return fail(BEVGEN_ERR_ARG, "dec_attention: bad args");
  CHECK_LAUNCH(launch_dec_attn(qkv_partials, ks, zstride, qkv_bias, y, camera_bias, bias_ld, k_cache, v_cache, kv_bf16, x1, step_ptr, workspace,
                               counters, batch, n_cond, heads, d, lmax, scale, row_counters, ln_gamma, ln_beta, ln_eps, (uint16_t*)ln_hi,
                               (uint16_t*)ln_lo, layout, layout_block, layout_ld, g_sm_count, (cudaStream_t)stream), "dec_attention");
}

BEVGEN_API int bevgen_dec_attention_workspace_floats(int batch, int heads) { return dec_attn_workspace_floats(batch, heads); }

BEVGEN_API int bevgen_sample_topk(const float* logit_partials, int ks, long long zstride, int vpad, int vocab, float temperature, int top_k,
                                  int greedy, unsigned long long seed, const long long* forced_tokens, const int* forward_shuffle_idx,
                                  long long* cam_idx, long long* tokens_out, float* logits_trace, float* probs_out, const int* step_ptr, int batch,
                                  int n_img, int hw, int ncam, void* stream) {
  if (!logit_partials || !forward_shuffle_idx || !cam_idx || !step_ptr || ks < 1) return fail(BEVGEN_ERR_ARG, "sample_topk: bad args");
  CHECK_LAUNCH(launch_dec_sample(logit_partials, ks, zstride, vpad, vocab, temperature, top_k, greedy, seed, forced_tokens, forward_shuffle_idx,
                                 cam_idx, tokens_out, logits_trace, probs_out, step_ptr, batch, n_img, hw, ncam, (cudaStream_t)stream), "sample_topk");
}

BEVGEN_API int bevgen_dec_advance(int* step_ptr, void* stream) {
  if (!step_ptr) return fail(BEVGEN_ERR_ARG, "dec_advance: null step");
  CHECK_LAUNCH(launch_dec_advance(step_ptr, (cudaStream_t)stream), "dec_advance");
}

BEVGEN_API int bevgen_decode_workspace(int batch, int d, int heads, int vocab, long long* n_floats, long long* n_counters) {
  if (!n_floats || !n_counters || batch < 1 || d < 64 || heads < 1 || vocab < 1) return fail(BEVGEN_ERR_ARG, "decode_workspace: bad argument");
  decode_workspace_sizes(batch, d, heads, vocab, n_floats, n_counters);
  return BEVGEN_OK;
}

BEVGEN_API long long bevgen_pack_decode_linear(const float* w, int n_rows, int ld, int d, int n_quarters, float lo_mul, void* out, void* stream) {
  if (n_rows < 1 || d < 64 || d % 64 != 0 || n_quarters < 1) return fail(BEVGEN_ERR_ARG, "pack_decode_linear: bad shape");
  const long long bytes = decode_packed_bytes(n_rows, d, n_quarters);
  if (out == nullptr) return bytes;
  if (!w) return fail(BEVGEN_ERR_ARG, "pack_decode_linear: null weight");
  const int rc = launch_pack_decode_linear(w, n_rows, ld, d, n_quarters, lo_mul, out, (cudaStream_t)stream);
  if (rc != BEVGEN_OK) return fail(rc, "pack_decode_linear: launch failed");
  return bytes;
}

BEVGEN_API int bevgen_decode_persistent(const bevgen_decode_args* a, void* stream) {
  if (!a || !a->layers || !a->w_head || !a->c1_head || !a->c2_head || !a->cam_idx || !a->x_tok_emb || !a->x_pos_emb || !a->forward_shuffle_idx ||
      !a->workspace || !a->counters)
    return fail(BEVGEN_ERR_ARG, "decode_persistent: null argument");
  if (a->img_embed_w && (!a->cam_embed_w || !a->intrinsics_inv || !a->extrinsics_inv || !a->pixel))
    return fail(BEVGEN_ERR_ARG, "decode_persistent: ray embedding inputs missing");
  if (g_sm_count < 16) return fail(BEVGEN_ERR_ARCH, "decode_persistent: needs at least 16 SMs");
  static_assert(sizeof(bevgen_decode_layer) == sizeof(DecodeLayer), "bevgen_decode_layer and DecodeLayer must have the same layout");
  DecodeParams p = {};
  p.layers = reinterpret_cast<const DecodeLayer*>(a->layers);
  p.n_layers = a->n_layers;
  p.w_head = (const uint8_t*)a->w_head; p.s_head = a->s_head; p.c1_head = a->c1_head; p.c2_head = a->c2_head;
  p.B = a->batch; p.d = a->d; p.H = a->heads; p.vocab = a->vocab; p.nc = a->n_cond; p.n_img = a->n_img; p.Lmax = a->lmax; p.ncam = a->ncam; p.hw = a->hw;
  p.step_begin = a->step_begin; p.step_end = a->step_end;
  p.cam_idx = a->cam_idx; p.x_tok_emb = a->x_tok_emb; p.x_pos_emb = a->x_pos_emb; p.img_embed_w = a->img_embed_w; p.cam_embed_w = a->cam_embed_w;
  p.I_inv = a->intrinsics_inv; p.E_inv = a->extrinsics_inv; p.pixel = a->pixel; p.fwd = a->forward_shuffle_idx;
  p.bias = a->camera_bias; p.bias_ld = a->bias_ld; p.scale = a->scale; p.temperature = a->temperature;
  p.top_k = a->top_k; p.greedy = a->greedy; p.seed = a->seed; p.forced = a->forced_tokens; p.tokens_out = a->tokens_out; p.trace = a->logits_trace;
  p.lay_blk = a->layout_block; p.lay_ld = a->layout_ld;
  p.debug = a->debug; p.profile = a->profile;
  CHECK_LAUNCH(launch_decode_persistent(p, a->workspace, a->counters, g_sm_count, (cudaStream_t)stream), "decode_persistent");
}

}  // extern "C"
