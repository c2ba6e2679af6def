// Internal launcher interface shared by the kernel translation units and c_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_tc.cuh"

namespace bevgen {

// Launch helper for the kernels of the KV-cache decode chain: with g_pdl_enabled (bevgen_set_pdl) the launch carries the programmatic
// stream-serialization attribute, so the kernel's prologue (barrier init, TMEM allocation, KV-slab prefetch) overlaps the tail of its
// predecessor; the kernels call pdl_wait() before touching dependent memory.  Works under stream capture (programmatic graph edges).
extern int g_pdl_enabled;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

enum PrepMode : int { PREP_IDENT = 0, PREP_UP2 = 1, PREP_S2D = 2 };

struct PrepParams {
  const float* x;          // [N][H][W][C]
  const float* mean_rstd;  // [N][32][2] (mean, rstd) or null (no normalisation)
  const float* gamma;      // [C]
  const float* beta;       // [C]
  uint16_t* hi;            // bf16 planes
  uint16_t* lo;            // may be null (single-pass bf16 mode)
  int N, H, W, C;          // source geometry
  int mode;                // PrepMode
  int swish;               // apply x*sigmoid(x) after the affine
};

struct EmbedParams {
  const long long* cam_idx;    // [B][ncam][hw]
  const long long* bev_idx;    // [B][nc]
  const float* I_inv;          // [B][ncam][3][3]
  const float* E_inv;          // [B][ncam][4][4]
  const float* x_tok_emb;      // [vocab+1][d]
  const float* cond_tok_emb;   // [cond_vocab][d]
  const float* x_pos_emb;      // [n_img][d]
  const float* cond_static;    // [nc][d]
  const float* img_embed_w;    // [d][4] or null
  const float* cam_embed_w;    // [d][4] or null
  const int* fwd;              // [n_img]
  const float* pixel;          // [hw][3]
  float* out;                  // [B][nrows][d]
  const int* step_ptr;         // decode: first row = nc + *step_ptr - 1 (overrides row0), or null
  int B, ncam, hw, nc, n_img, L, d, vocab;
  int pad_last;                // teacher forcing: the last (cam,h,w) token is replaced by PAD (:328-329)
  int bev_embed;               // subtract sum_cam c_embed on cond rows
  int row0, nrows;             // sequence rows [row0, row0+nrows) are produced; out row index = s - row0
};

int gemm_tc_dispatch(const GemmParams& p, int bn, int npass, int sm_count, cudaStream_t stream);

int launch_gn_stats(const float* x, double* sums, float* mean_rstd, int N, int pixels, int C, float eps, cudaStream_t st);
int launch_prep(const PrepParams& p, int sm_count, cudaStream_t st);
int launch_gn_finalize(const double* sums, float* mean_rstd, int N, int pixels, int C, float eps, cudaStream_t st);
int launch_im2col3x3(const float* x, uint16_t* hi, uint16_t* lo, int N, int Cin, int H, int W, int sm_count, cudaStream_t st);
int launch_transpose(const float* src, float* dst, int N, int R, int Cc, cudaStream_t st);
int launch_softmax_rows(const float* s, uint16_t* hi, uint16_t* lo, long long rows, int cols, int out_ld, float scale, cudaStream_t st);
int launch_gather_rows(const float* table, const long long* idx, float* out, long long rows, int D, int n_table, int sm_count, cudaStream_t st);
int launch_conv_in3(const float* x, const float* w, const float* bias, float* out, double* gn_sums, int N, int H, int W, int Cout, int sm_count,
                    cudaStream_t st);
int launch_conv_out3(const float* x, const float* affine, int swish, const float* w, const float* bias, float* out, int N, int H, int W, int C,
                     int sm_count, cudaStream_t st);
int launch_absmax(const float* x, long long n, float* out, int sm_count, cudaStream_t st);
int launch_split_bf16(const float* x, long long n, void* hi, void* lo, int sm_count, cudaStream_t st);
int launch_pack_f16f8(const float* w, long long rows, int cin, int chunk, float s, float w16_mul, void* w16, void* pair, int sm_count, cudaStream_t st);
int launch_to_uint8_hwc(const float* x, uint8_t* out, int N, int C, int P, int sm_count, cudaStream_t st);
int launch_denorm(const float* x, float* out, int N, int C, int P, const float* mean, const float* std_, int sm_count, cudaStream_t st);
int launch_row_sqnorm(const float* x, float* out, int rows, int D, cudaStream_t st);
int launch_vq_nearest(const float* z, const float* cb, const float* zz, const float* ee, long long* idx, float* zq, int rows, int n_codes,
                      int D, cudaStream_t st);

int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, uint16_t* hi, uint16_t* lo, long long rows, int d,
                     long long x_row_stride, float eps, int f16f8, cudaStream_t st);
int launch_embed(const EmbedParams& p, cudaStream_t st);
int launch_attn_softmax(const float* S, const float* bias, const uint8_t* mask, uint16_t* hi, uint16_t* lo, long long zrows, int L, int Lk,
                        float scale, const uint8_t* layout, int H, int blk, int lay_ld, cudaStream_t st);

struct ConvHaloParams {
  CUtensorMap tmA[2];      // 5D (c8, w, h, chunk, n), box (8, 10, 18, 8, 1), no swizzle
  CUtensorMap tmW[2];      // 2D (cin, tap*cout + co), box (64, 128), SWIZZLE_128B
  int N, H, W, Cin, Cout;
  const float* bias;
  const float* residual;   // [N][H][W][Cout] or null
  float* out;              // [N][H][W][Cout]
  double* gn_sums;         // [N][32][2] or null: GroupNorm statistics of `out`
};

int launch_conv_halo(const ConvHaloParams& p, int npass, int sm_count, cudaStream_t st);

struct ConvFusedParams {
  CUtensorMap tmW[2];      // 2D (cin, tap*cout + co), box (64, 128), SWIZZLE_128B
  const float* x;          // fp32 NHWC source [N][Hs][Ws][Cin] (Hs = H/2, Ws = W/2 when up2)
  const float* affine;     // [N][Cin][2] (scale, shift) of the fused GroupNorm, or null
  int swish, up2;
  int N, H, W, Cin, Cout;  // H, W: conv (= output) geometry
  const float* bias;
  const float* residual;   // [N][H][W][Cout] or null
  float* out;              // [N][H][W][Cout]
  double* gn_sums;         // [N][32][2] or null: GroupNorm statistics of `out`
  int trace_cta, trace_step;       // BEVGEN_DP_DBG & 64: thread 0 of this CTA records a clock trace of one layer of this step behind the profile rows
  int dbg;                 // ablation switches (BEVGEN_CONV_DBG, tools/conv_ablation.py): 1 no global fetch, 2 no operand transform/stores,
                           // 4 epilogue drains TMEM only, 8 no MMA issue, 16 no weight traffic.  0 in production.
  float lo_scale;          // npass == 2 (fp16 + e4m3 corrections): 1 / (2^13 * weight scale), applied to the correction accumulator
};
int launch_conv_fused(const ConvFusedParams& p, int npass, int sm_count, cudaStream_t st);
int launch_conv_fused2(const ConvFusedParams& p, int npass, int sm_count, cudaStream_t st);   // 2-CTA clusters; tmW box = (64, 64)
int launch_conv_fused3(const ConvFusedParams& p, int npass, int sm_count, cudaStream_t st);   // 2-CTA clusters, 16x16 blocks; tmW box = (32, 64), SWIZZLE_64B
int launch_gn_affine(const double* sums, const float* gamma, const float* beta, float* affine, int N, int pixels, int C, float eps, cudaStream_t st);

struct GemmPairParams {
  CUtensorMap tmA[2];      // A16s / Apair: 2D (k, row), box (64, 128), SWIZZLE_128B
  CUtensorMap tmW[2];      // W16s / Wpair: 2D (k, out feature), box (64, 128)
  int M, N, K;
  float out_scale;         // 1 / (2^13 * S)
  const float* bias;       // [N] or null
  int gelu;
  const float* residual;   // [M][N] or null
  float* out_f32;          // [M][N] or null
  uint16_t* out_hi;        // bf16 hi / lo planes [M][N] or null (both)
  uint16_t* out_lo;
  uint16_t* out_f16;       // scaled f16f8 planes for a following GEMM: fp16(y * 2^6) [M][N] + e4m3 pair plane [M][2N bytes], or null (both)
  void* out_pair;
};
int launch_gemm_pair_f16f8(const GemmPairParams& p, int sm_count, cudaStream_t st);

int launch_ray_embed_add(float* h, const float* I_inv, const float* E_inv, const float* pixel, const float* img_w, const float* cam_w, int n_images,
                         int hw, int d, cudaStream_t st);
int launch_mg_head_planes(const float* src, long long src_ld, int src_col0, int n_src, int src_batch_rows, const float* null_vec, const float* scale,
                          uint16_t* hi, uint16_t* lo, int B, int dst_rows, int dst_batch_rows, long long dst_ld, int dst_col0, int has_null, int H,
                          cudaStream_t st);
int launch_mg_sample(const float* logits, const float* u, long long* ids, float* scores, long long rows, int V, int k, float inv_temp, long long mask_id,
                     cudaStream_t st);
int launch_mg_remask(const float* scores, const float* u, float noise_scale, long long* ids, const long long* init_ids, long long rows, int hw, int n_mask,
                     long long mask_id, cudaStream_t st);
int launch_mg_geglu_ln(const float* hin, long long h_ld, const float* gamma, uint16_t* hi, uint16_t* lo, long long rows, int f, int f_pad, float eps,
                       int f16f8, cudaStream_t st);
int launch_attn_fused(const CUtensorMap* tm_hi, const CUtensorMap* tm_lo, const void* bias_f16, const float* y, float* x1, int B, int H, int L,
                      int nc, int d, float scale, int npass, const unsigned long long* layout64, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st);

int launch_dec_reduce_ln(const float* partials, int ks, long long zstride, const float* bias, const float* residual, long long res_stride,
                         const float* gamma, const float* beta, float eps, float* x_out, float* y, uint16_t* hi, uint16_t* lo, int rows, int d,
                         cudaStream_t st);
int launch_dec_reduce_act(const float* partials, int ks, long long zstride, const float* bias, uint16_t* hi, uint16_t* lo, int rows, int n,
                          int gelu, cudaStream_t st);
int launch_kv_store(const uint16_t* hi, const uint16_t* lo, void* kc, void* vc, int kv_bf16, int B, int Lp, int nrows, int H, int d, int Lmax,
                    cudaStream_t st);
int launch_dec_attn(const float* qkv_part, int ks, long long zstride, const float* bqkv, const float* y, const float* bias, int bias_ld,
                    void* kc, void* vc, int kv_bf16, float* x1, const int* step_ptr, float* ws, unsigned int* counters, int B, int nc, int H,
                    int d, int Lmax, float scale, unsigned int* row_counters, const float* ln_gamma, const float* ln_beta, float ln_eps,
                    uint16_t* ln_hi, uint16_t* ln_lo, const uint8_t* layout, int lay_blk, int lay_ld, int sm_count, cudaStream_t st);
int dec_attn_workspace_floats(int B, int H);
int launch_dec_sample(const float* part, int ks, long long zstride, int vpad, int V, float temperature, int top_k, int greedy,
                      unsigned long long seed, const long long* forced, const int* fwd, long long* cam_idx, long long* tokens_out, float* trace,
                      float* probs_out, const int* step_ptr, int B, int n_img, int hw, int ncam, cudaStream_t st);
int launch_dec_advance(int* step, cudaStream_t st);

// ---------------------------------------------------------------- persistent KV-cache decode kernel (decode_persistent.cu)
struct DecodeLayer {               // one per transformer block, array in DEVICE memory (mirrors bevgen_decode_layer)
  const uint8_t* w_qkv;            // packed by launch_pack_decode_linear from gamma1-scaled [3d][d]: [3d/8 units][d/64][1536 B]
  const uint8_t* w_1;              // gamma2-scaled mlp.0.weight: [4d/8 units]
  const uint8_t* w_2;              // mlp.2.weight: [4 K-quarters][d/8 units]
  const float* c1_qkv;             // [3d] sum_k gamma1_k W_nk           (lazy LayerNorm: out = rstd * (acc - mean * c1) + c2)
  const float* c2_qkv;             // [3d] bias_n + sum_k beta1_k W_nk
  const float* c2_2;               // [d]  mlp.2.bias
  const float* ln1_g;              // [d]  LayerNorm 1 (the residual is taken from its output)
  const float* ln1_b;
  const float* c1_1;               // [4d] sum_k gamma2_k W1_nk
  const float* c2_1;               // [4d] bias1_n + sum_k beta2_k W1_nk
  void* kc;                        // fp16 K cache [B][H][Lmax/128][64][128]
  void* vc;                        // fp16 V cache [B][H][Lmax][64]
  const uint8_t* layout;           // optional per-head block layout [H][lay_ld][lay_ld]
  float s_qkv, s_1, s_2, pad_;     // 1 / S of the e4m3 residual planes
};
struct DecodeParams {
  const DecodeLayer* layers;
  int n_layers;
  const uint8_t* w_head;           // gamma_f-scaled head.weight
  float s_head;
  const float *c1_head, *c2_head;  // [vpad] lazy ln_f constants
  int B, d, H, vocab, vpad, nc, n_img, Lmax, ncam, hw;
  int step_begin, step_end;        // decode-order tokens [step_begin, step_end) are produced; token step_begin - 1 is already in cam_idx
  long long* cam_idx;              // [B][ncam][hw] token grid (read for the first embedding, written per step)
  const float *x_tok_emb, *x_pos_emb, *img_embed_w, *cam_embed_w, *I_inv, *E_inv, *pixel;
  const int* fwd;                  // forward_shuffle_idx
  const float* bias;               // camera bias [L][bias_ld] (may be NULL)
  int bias_ld;
  float scale, temperature;
  int top_k, greedy;
  unsigned long long seed;
  const long long* forced;         // [B][n_img] or NULL
  long long* tokens_out;           // [B][n_img] or NULL
  float* trace;                    // [n_img][B][vocab] or NULL
  int lay_blk, lay_ld;
  // workspace (filled by the launcher)
  float *X, *X1, *QKV, *LOGITS, *PSX, *PSX1;
  uint8_t *XF, *X1F, *HF;          // activation vectors in mma A-fragment order (fp16 hi + lo)
  float* P2;                       // MLP2 partial sums per K-quarter
  unsigned int* barrier;
  int trace_cta, trace_step;       // BEVGEN_DP_DBG & 64: thread 0 of this CTA records a clock trace of one layer of this step behind the profile rows
  int dbg;                         // timing experiments (BEVGEN_DP_DBG): 1 skip attention math, 2 skip linear MMAs, 4 skip activation fetches, 8 producer copies nothing
  unsigned int* debug;             // optional pinned HOST buffer (8 uint32, zeroed): timeout diagnostics written before the trap
  unsigned long long* profile;     // optional [grid][32] nanoseconds per phase (bodies and grid barriers), device memory
};
int launch_decode_persistent(DecodeParams p, float* ws, unsigned int* counters, int sm_count, cudaStream_t st);
int launch_pack_decode_linear(const float* W, int n_rows, int ld, int d, int n_quarters, float lo_mul, void* out, cudaStream_t st);
long long decode_packed_bytes(int n_rows, int d, int n_quarters);
void decode_workspace_sizes(int B, int d, int H, int vocab, long long* n_floats, long long* n_counters);


}  // namespace bevgen
