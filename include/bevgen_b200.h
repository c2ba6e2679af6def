/* bevgen_b200 — C ABI of the B200 (sm_100a) kernels behind BEVGen's two hot paths.
 *
 * The reference (alexanderswerdlow/BEVGen) is pure Python: its "operator API" for these paths is a set of
 * torch.nn.Module calls (cuDNN/cuBLAS/DeepSpeed-Triton underneath).  Each entry point below names the reference
 * call sites it replaces (paths relative to /root/reference/multi_view_generation).  Conventions:
 *   - plain device pointers + sizes, a CUDA stream passed as void* (cudaStream_t); no allocation, no host sync;
 *   - activations are NHWC; "bf16 planes" are the (hi, lo) split of an fp32 tensor (x ~= hi + lo), lo may be NULL
 *     for single-pass bf16 arithmetic (npass = 1); npass = 3 is the fp32-equivalent bf16x3 product;
 *   - every function returns 0 on success or a negative BEVGEN_ERR_* code; bevgen_last_error() has the text.
 * INTEGRATION.md shows the ctypes binding used on the Python side.
 */
#ifndef BEVGEN_B200_H
#define BEVGEN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BEVGEN_API __attribute__((visibility("default")))
#else
#define BEVGEN_API
#endif

#define BEVGEN_OK 0
#define BEVGEN_ERR_ARG (-1)    /* bad shape / alignment / unsupported size */
#define BEVGEN_ERR_CUDA (-2)   /* CUDA runtime error (launch failure, ...) */
#define BEVGEN_ERR_ARCH (-3)   /* device is not sm_100 */
#define BEVGEN_ERR_DRIVER (-4) /* driver entry point (cuTensorMapEncodeTiled) unavailable */

#define BEVGEN_MAX_TAPS 9

/* flags of bevgen_gemm_args.flags */
#define BEVGEN_GF_GELU 1        /* exact-erf GELU after bias (mingpt_sparse.py:235) */
#define BEVGEN_GF_OUT_NCHW 2    /* fp32 output stored [z][col][h][w] (tiny Cout, e.g. conv_out -> NCHW images) */
#define BEVGEN_GF_B_MN 4        /* B operand is [k rows][n cols] in memory (V in P.V) */
#define BEVGEN_GF_CAUSAL_SKIP 8 /* skip output tiles outside the [cond | causal] support (Q.K^T) */
#define BEVGEN_GF_OUT_T 32 /* store D^T: out_f32[z][col][row] (swap-AB decode GEMMs, split-K partials; no bias/act/residual) */
#define BEVGEN_GF_OUT_F16F8 64 /* out_hi / out_lo receive the fp16 plane and the e4m3 pair plane (operands of a following npass = 2 GEMM) */
#define BEVGEN_GF_CAUSAL_KLIMIT 16 /* reduction over keys stops at max(ncond, last tile row + 1) (P.V) */

/* prep modes */
#define BEVGEN_PREP_IDENT 0
#define BEVGEN_PREP_UP2 1 /* nearest 2x upsample (model.py:49-53) folded into the operand write */
#define BEVGEN_PREP_S2D 2 /* space-to-depth phase planes for the stride-2 asymmetric-pad conv (model.py:68-75) */

/* One implicit-GEMM launch: D[z][pixel][col] = sum_tap sum_k A[z, pixel + tap offset][k] * B[tap][col][k] (+bias, act, residual).
 * Serves torch.nn.Conv2d 3x3/1x1 (stage1/model.py:43-47,62-66,88-115,146-165,355-359,399-403,458-462,500-504; vqgan.py:72-73),
 * nn.Linear (mingpt_sparse.py:173-175,233-238,286) and the batched attention products (model.py:178-187;
 * sparse_self_attention.py:153,176). */
typedef struct {
  const void* a_hi; const void* a_lo;   /* bf16 planes, logical [a_n][a_h][a_w][a_c] */
  int a_n, a_h, a_w, a_c;
  const void* b_hi; const void* b_lo;   /* bf16 planes, row-major [b_rows][b_cols] */
  int b_rows, b_cols;
  int ntaps;
  int tap_dx[BEVGEN_MAX_TAPS], tap_dy[BEVGEN_MAX_TAPS], tap_dn[BEVGEN_MAX_TAPS];
  int a_n_mul, a_n_zstride;             /* A image index = z_outer*a_n_mul + z_inner*a_n_zstride + tap_dn[tap] */
  int k;                                /* reduction length per tap, multiple of 64 */
  int a_c_off, a_c_zstride;             /* A channel offset = a_c_off + z_inner*a_c_zstride */
  int b_k_off, b_k_zstride;             /* B k offset (K-major) or column offset (MN-major) */
  int b_row_zstride, b_row_tapstride;   /* B row = z_outer*b_row_zstride + tap*b_row_tapstride + col */
  int z_inner, z_outer;
  int tile_w, tile_h;                   /* output tile, tile_w*tile_h == 128 */
  int out_w, out_h, n_cols;
  long long out_zo_stride, out_zi_stride;
  int ldc;
  const float* bias; const float* residual;
  float* out_f32; void* out_hi; void* out_lo;
  int flags;
  int causal_ncond;
  int bn;                               /* N tile: 16, 64 or 128 */
  int npass;                            /* 1 (bf16), 3 (bf16x3 ~ fp32) or 2 (f16f8 ~ fp32: a_hi / b_hi are fp16 planes, a_lo / b_lo the e4m3 pair planes
                                           of ops.pack_f16f8 / BEVGEN_GF_OUT_F16F8 / bevgen_layernorm_f16f8; K-major B, bn = 128 only) */
  /* split-K finalize fused into BEVGEN_GF_OUT_T launches: the last CTA per 128-feature tile reduces the z_inner partials */
  int fin_mode;                         /* 0 none; 1: planes = act(sum + bias); 2: x = sum + bias + residual, then LayerNorm(x) */
  int fin_gelu, fin_rows;
  const float* fin_bias; const float* fin_resid;
  float* fin_x; float* fin_y; void* fin_hi; void* fin_lo;
  const float* fin_gamma; const float* fin_beta;
  float fin_eps;
  unsigned int* fin_counters;           /* [ceil(out_w/128) + 1] uint32, zeroed once by the caller (self-resetting) */
  float lo_scale;                       /* npass == 2: 1 / (2^13 * S), S = weight scale (ops.pack_f16f8) */
} bevgen_gemm_args;

BEVGEN_API int bevgen_init(int device);                 /* selects nothing, queries: SM count, arch check, driver entry points */
BEVGEN_API const char* bevgen_last_error(void);
BEVGEN_API int bevgen_version(void);
BEVGEN_API int bevgen_sm_count(void);
/* Programmatic dependent launch for the KV-cache decode chain (gemm_tc, dec_*, embed, sample kernels): when enabled, those kernels are
 * launched with the programmatic stream-serialization attribute and overlap their prologue with the previous kernel's tail (they call
 * griddepcontrol.wait before touching dependent memory).  Returns the previous setting.  Process-wide; the sampler brackets its steps. */
BEVGEN_API int bevgen_set_pdl(int enabled);

BEVGEN_API int bevgen_gemm_tc(const bevgen_gemm_args* args, void* stream);

/* 3x3 stride-1 "same" Conv2d (stage1/model.py:43-47,88-102,355-359,399-403,458-462,500-504) over NHWC bf16 operand planes with a
 * shared-memory halo tile (activations loaded once per 64-channel chunk instead of once per tap), fused bias + residual and, when
 * gn_sums != NULL, the GroupNorm(32) statistics of the OUTPUT: gn_sums[n][32][2] (sum, sum of squares; zeroed by this call), to be
 * turned into (mean, rstd) by bevgen_groupnorm_finalize.  Weights: [tap][cout][cin] planes, rows padded to 8*cout + ceil128(cout). */
BEVGEN_API int bevgen_conv3x3_halo(const void* a_hi, const void* a_lo, int n, int h, int w, int cin, const void* w_hi, const void* w_lo, int w_rows,
                                   int cout, const float* bias, const float* residual, float* out, double* gn_sums, int npass, void* stream);
/* Same convolution reading the fp32 NHWC activation directly: GroupNorm-apply (per-(image,channel) affine from bevgen_groupnorm_affine,
 * or NULL) + swish (model.py:29-31) + bf16 split + optional nearest 2x upsampling (up2: x is [n][h/2][w/2][cin], model.py:49-53) are
 * fused into the operand path of the conv (no `prep` pass, no operand planes in HBM).  npass | 0x100 selects the cta_group::2 kernel
 * (clusters of two CTAs sharing every weight tile: half the L2->SM weight traffic and shared-memory reads per MMA). */
BEVGEN_API int bevgen_conv3x3_fused(const float* x, int n, int h, int w, int cin, const float* affine, int swish, int up2, const void* w_hi,
                                    const void* w_lo, int w_rows, int cout, const float* bias, const float* residual, float* out, double* gn_sums,
                                    int npass, void* stream);
/* Same convolution (cta_group::2 kernels) with the fp32-equivalent product formed as ONE fp16 MMA plus TWO e4m3 MMAs (which run at twice
 * the fp16 rate) instead of three bf16 MMAs.  Max error ~2^-15 relative per product (bf16x3: ~2^-17).
 * block16 == 0 (conv_fused2.cu, any input range):
 *     x*w ~= x16*w16 + [ e4m3((x - x16) * 2^13) * e4m3(w * S)  +  e4m3(x) * e4m3((w - w16) * S * 2^13) ] * lo_scale,   lo_scale = 1 / (2^13 * S)
 *   the correction sum has its own TMEM accumulator; out-of-range values saturate (the element degrades to fp16 accuracy, no overflow).
 *     w_f16    [tap][cout][cin] fp16, rows padded like bevgen_conv3x3_fused
 *     w_f8pair same rows, 2*cin bytes per row: per 64-channel chunk 64 bytes e4m3(w * S) followed by 64 bytes e4m3((w - w16) * S * 2^13)
 * block16 == 1 (conv_fused3.cu, weight-stationary 16x16 pixel blocks for the large feature maps; |x| < 1024 required, i.e. GroupNorm-ed inputs):
 *     x*w ~= [ fp16(x * 2^6) * w16s + e4m3((x - x16) * 2^13) * e4m3(w * S) + e4m3(x) * e4m3((w - w16) * S * 2^13) ] * lo_scale
 *   one accumulator; w_f16 = fp16(w * S * 2^7) (so w16 = w_f16 / (S * 2^7)), and the pair rows hold, per 32-channel slice, 32 bytes
 *   e4m3(w * S) followed by 32 bytes e4m3((w - w16) * S * 2^13).  cin % 32 == 0.
 * bevgen_conv3x3_fused with npass | 0x200 selects the block kernel for the bf16 / bf16x3 products (same weight planes). */
BEVGEN_API int bevgen_conv3x3_fused_f16f8(const float* x, int n, int h, int w, int cin, const float* affine, int swish, int up2, const void* w_f16,
                                          const void* w_f8pair, int w_rows, int cout, float lo_scale, const float* bias, const float* residual,
                                          float* out, double* gn_sums, int block16, void* stream);
/* (sum, sumsq) per (image, group) + GroupNorm weight/bias -> affine[n][c][2] = (rstd*gamma, beta - mean*rstd*gamma) */
BEVGEN_API int bevgen_groupnorm_affine(const double* sums, const float* gamma, const float* beta, int n, int pixels, int c, float eps, float* affine,
                                       void* stream);
BEVGEN_API int bevgen_groupnorm_finalize(const double* sums, int n, int pixels, int c, float eps, float* mean_rstd, void* stream);

/* torch.nn.GroupNorm(32, C, eps) statistics (stage1/model.py:34-35): fp32 NHWC x[n][pixels][c] -> mean_rstd[n][32][2].
 * ws_sums: n*64 doubles of scratch. */
BEVGEN_API int bevgen_groupnorm_stats(const float* x, int n, int pixels, int c, float eps, double* ws_sums, float* mean_rstd, void* stream);

/* GroupNorm-apply (+ swish, model.py:29-31) + bf16 split + optional spatial remap; mean_rstd NULL = no normalisation. */
BEVGEN_API int bevgen_prep_operand(const float* x, int n, int h, int w, int c, const float* mean_rstd, const float* gamma, const float* beta,
                        int swish, int mode, void* out_hi, void* out_lo, void* stream);

/* conv_in operand: fp32 NCHW image (cin*9 <= 64) -> [n][h][w][64] bf16 planes, k = (kh*3+kw)*cin + c (model.py:355-359). */
BEVGEN_API int bevgen_im2col3x3(const float* x_nchw, int n, int cin, int h, int w, void* out_hi, void* out_lo, void* stream);

/* fp32 [n][r][c] -> [n][c][r] */
BEVGEN_API int bevgen_transpose_f32(const float* src, float* dst, int n, int r, int c, void* stream);

/* softmax(scale * s) per row of fp32 [rows][cols] -> bf16 planes with row pitch out_ld >= cols (model.py:179-180) */
BEVGEN_API int bevgen_softmax_rows(const float* s, long long rows, int cols, float scale, void* out_hi, void* out_lo, int out_ld, void* stream);

/* VectorQuantizer2.forward (stage1/quantize.py:276-285): idx[r] = argmin_j |z_r|^2 + |e_j|^2 - 2 z_r.e_j ; zq = e[idx] (optional).
 * code_sqnorm[n_codes] from bevgen_row_sqnorm(codebook); ws_zz: rows floats of scratch. */
BEVGEN_API int bevgen_row_sqnorm(const float* x, int rows, int dim, float* out, void* stream);
BEVGEN_API int bevgen_vq_nearest(const float* z, const float* codebook, const float* code_sqnorm, int rows, int n_codes, int dim, float* ws_zz,
                      long long* idx, float* zq, void* stream);

/* VectorQuantizer2.get_codebook_entry (quantize.py:314-329), NHWC result: out[r][:] = codebook[idx[r]][:] */
BEVGEN_API int bevgen_codebook_gather(const float* codebook, const long long* idx, long long rows, int dim, int n_codes, float* out, void* stream);

/* bev_utils/util.py:97-118 denormalize_tensor(keep_tensor=True) on fp32 NCHW, 3 channels */
BEVGEN_API int bevgen_denormalize(const float* x, float* out, int n, int c, int pixels, const float* mean3, const float* std3, void* stream);
/* Encoder.conv_in for RGB inputs (stage1/model.py:355-359): 3x3 "same" conv fp32 NCHW [n][3][h][w] -> fp32 NHWC [n][h][w][cout],
 * cout = 64 or 128, weights OIHW fp32 as stored in the checkpoint, exact fp32 FMA on the CUDA cores (the implicit-GEMM form pads K = 27
 * to 64 and needs a 1.6 GB im2col plane per 96 images).  gn_sums (may be NULL): GroupNorm(32) statistics of the output [n][32][2]. */
BEVGEN_API int bevgen_conv_in3(const float* x_nchw, const float* weight_oihw, const float* bias, float* out_nhwc, double* gn_sums, int n, int h, int w,
                               int cout, void* stream);
/* Decoder.norm_out -> swish -> conv_out for RGB outputs (stage1/model.py:500-504,533-536): fp32 NHWC [n][h][w][c] (c = 64 or 128) ->
 * fp32 NCHW [n][3][h][w]; affine = per-(image, channel) GroupNorm (scale, shift) from bevgen_groupnorm_affine or NULL, swish applied
 * when nonzero; weights OIHW [3][c][3][3] fp32; exact fp32 FMA on the CUDA cores, the input is read once. */
BEVGEN_API int bevgen_conv_out3(const float* x_nhwc, int n, int h, int w, int c, const float* affine, int swish, const float* weight_oihw,
                                const float* bias, float* out_nchw, void* stream);
/* fp32 NCHW images in [0,1] -> uint8 NHWC, round to nearest (the device half of GenerateImages.save_raw_data / save_img,
 * utils/callback.py:28-30,72-132: what leaves the GPU is a quarter of the fp32 bytes, already in the layout the JPEG encoder wants). */
BEVGEN_API int bevgen_to_uint8_hwc(const float* x_nchw, void* out_nhwc_u8, int n, int c, int pixels, void* stream);

/* ---------------------------------------------------------------- weight packing (model load; SURVEY 8b "bevgen_pack_*")
 * The reference keeps nn.Parameter tensors as they are (stage1/model.py, transformer/mingpt_sparse.py); this library's kernels read
 * pre-split operand planes, written once per weight version by these three entry points (bevgen_pack_decode_linear below packs the
 * persistent decode kernel's slabs). */
/* max |x| over n floats; *out must be 0.0f before the call */
BEVGEN_API int bevgen_absmax(const float* x, long long n, float* out, void* stream);
/* fp32 -> bf16 hi and (lo may be NULL) bf16(x - hi) planes: the bf16x3 split-product operand */
BEVGEN_API int bevgen_pack_split_bf16(const float* x, long long n, void* hi_bf16, void* lo_bf16, void* stream);
/* fp32 weight rows [rows][cin] -> f16f8 operand: w16 = fp16(w * w16_mul) [rows][cin]; pair [rows][2 * cin] bytes: per chunk (32 | 64)
 * elements `chunk` bytes e4m3(w * s) then `chunk` bytes e4m3((w - w16 / w16_mul) * s * 2^13).  s = the power of two with max|w| * s <= 64. */
BEVGEN_API int bevgen_pack_f16f8(const float* w, long long rows, int cin, int chunk, float s, float w16_mul, void* w16, void* pair_u8, void* stream);

/* ---------------------------------------------------------------- stage-2 transformer */

/* nn.LayerNorm(d) (mingpt_sparse.py:220-221,285): x fp32 rows (pitch x_row_stride elements) -> y fp32 [rows][d] (optional)
 * and bf16 planes [rows][d] (optional). d % 128 == 0, d <= 1024. */
BEVGEN_API int bevgen_layernorm(const float* x, long long rows, int d, long long x_row_stride, const float* gamma, const float* beta, float eps,
                                float* y, void* out_hi, void* out_lo, void* stream);
/* Same LayerNorm writing the operand planes of an f16f8 GEMM (bevgen_gemm_tc, npass = 2): out_f16 [rows][d] fp16 and out_f8pair
 * [rows][2*d bytes]: per 64-column chunk 64 bytes e4m3((y - fp16(y)) * 2^13) followed by 64 bytes e4m3(y). */
BEVGEN_API int bevgen_layernorm_f16f8(const float* x, long long rows, int d, long long x_row_stride, const float* gamma, const float* beta, float eps,
                                      float* y, void* out_f16, void* out_f8pair, int scaled, void* stream);
/* scaled != 0: the fp16 plane holds 2^6 * y (the e4m3 remainders stay relative to the unscaled fp16 value): A operand of bevgen_linear_f16f8. */

/* nn.Linear on the 2-CTA f16f8 GEMM (csrc/gemm_pair.cu; replaces the c_attn / mlp[0] / mlp[2] Linear calls of Block.forward,
 * mingpt_sparse.py:170-175,231-237 at fp32-equivalent precision): out[M][N] = act(A . W^T * out_scale + bias) (+ residual).
 * a16 [M][K] fp16(a * 2^6) and apair [M][2K bytes] come from bevgen_layernorm_f16f8(scaled = 1) or a previous call's out_f16 / out_pair;
 * w16 [N][K] = fp16(w * S * 2^7), wpair [N][2K bytes] = per 64-element chunk 64 B e4m3(w * S) then 64 B e4m3((w - w16) * S * 2^13),
 * out_scale = 1 / (2^13 * S) (ops.pack_linear_f16f8).  Needs |a| < 1024, N % 32 == 0, K % 64 == 0.
 * Outputs (any non-empty subset): out_f32 [M][N]; bf16 planes out_hi / out_lo [M][N]; scaled f16f8 planes out_f16 [M][N] + out_pair [M][2N]. */
BEVGEN_API int bevgen_linear_f16f8(const void* a16, const void* apair, const void* w16, const void* wpair, long long M, int N, int K, float out_scale,
                                   const float* bias, int gelu, const float* residual, float* out_f32, void* out_hi, void* out_lo, void* out_f16,
                                   void* out_pair, void* stream);

/* Input-embedding assembly of GPT.forward (mingpt_sparse.py:319-373) for sequence rows [row0, row0+nrows):
 * token + ray embedding (L2-normalised) + position embeddings, decode-order permutation, [cond | img | pad] concat. */
typedef struct {
  const long long* cam_idx;  /* [B][ncam][hw] int64 tokens */
  const long long* bev_idx;  /* [B][nc] */
  const float* intrinsics_inv; /* [B][ncam][3][3] */
  const float* extrinsics_inv; /* [B][ncam][4][4] */
  const float* x_tok_emb;    /* [vocab+1][d] */
  const float* cond_tok_emb; /* [cond_vocab][d] */
  const float* x_pos_emb;    /* [n_img][d] */
  const float* cond_static;  /* [nc][d]: cond_pos_emb (+ bev_embed(grid) - sum_cam bev_cam_pos_emb) */
  const float* img_embed_w;  /* [d][4] or NULL */
  const float* cam_embed_w;  /* [d][4] or NULL */
  const int* forward_shuffle_idx; /* [n_img] */
  const float* pixel;        /* [hw][3] image-plane grid */
  float* out;                /* [B][nrows][d] */
  const int* step_ptr;       /* decode: device step counter s, first row = nc + s - 1 (overrides row0); NULL otherwise */
  int B, ncam, hw, nc, n_img, L, d, vocab;
  int pad_last, bev_embed;
  int row0, nrows;
} bevgen_embed_args;
BEVGEN_API int bevgen_embed_assemble(const bevgen_embed_args* args, void* stream);

/* P = softmax_j(scale * (S + bias)) over mask != 0, zeros elsewhere (sparse_self_attention.py:153-173, dense form).
 * S fp32 [zrows][Lk] with zrows = batch*heads*L rows, bias fp32 [L][Lk] or NULL, mask uint8 [L][Lk]. */
BEVGEN_API int bevgen_attn_softmax(const float* s, const float* bias, const unsigned char* mask, long long zrows, int L, int Lk, float scale,
                                   void* out_hi, void* out_lo, const unsigned char* layout, int heads, int block, int layout_ld, void* stream);
/* layout (may be NULL): per-head block layout of DeepSpeed's SparsityConfig (sparse_self_attention.py:59-60; density < 1 configs),
 * uint8 [heads][layout_ld][layout_ld], nonzero = the (query block, key block) pair of `block` x `block` positions is attended; rows of s
 * are ordered [batch][head][query]. */

/* Fused attention forward (one kernel for sparse_self_attention.py:153-176 + the residual add of mingpt_sparse.py:250):
 *   x1[b,i,h*64:(h+1)*64] = y[...] + sum_j softmax_j(scale * (q_i.k_j + bias[i][j])) v_j,  allowed(i,j) = j < n_cond || (i >= n_cond && j <= i)
 * q/k/v are the column blocks [0,d), [d,2d), [2d,3d) of the fused qkv bf16 planes [batch][seq_len][3d].  bias_tiled (shared by batch
 * and heads, or NULL) is the camera bias already multiplied by scale * log2(e), fp16, in 128 x 128 tiles laid out for coalesced reads:
 *   bias_tiled[qt][kt][p][u][r][e] = fp16(scale * log2(e) * bias[128*qt + r][128*kt + 32*p + 8*u + e]),  p, u < 4, r < 128, e < 8
 * (`bevgen_b200.ops.tile_attention_bias`).  seq_len and n_cond must be multiples of 128, seq_len <= 4096, d = heads*64.
 * layout64 (NULL = dense [cond | causal] support): DeepSpeed block-sparse layouts of the density < 1 configs (sparse_self_attention.py:59-60,
 * mask_generator.py:217-228) at 16-position granularity, uint64 [heads][seq_len/128][seq_len/128]: bit (8*rb + kb) of entry [h][qt][kt] =
 * the head attends from query rows 128*qt + 16*rb .. +15 to keys 128*kt + 16*kb .. +15 (`ops.layout_to_tiles64`; sparse_block_size 16 (the
 * reference's configs/model/stage_2.yaml:22), 32, 64 or 128).
 * allowed(i,j) is ANDed with it; key tiles whose entry is 0 are skipped (no loads, no MMAs). */
BEVGEN_API int bevgen_attn_fused_fwd(const void* qkv_hi, const void* qkv_lo, int batch, int seq_len, int heads, int d, int n_cond,
                                     const void* bias_tiled, const float* y, float* x1, float scale, int npass, const unsigned long long* layout64,
                                     void* out_hi, void* out_lo, void* stream);
/* y may be NULL (no residual); the result goes to x1 (fp32) and / or to bf16 hi / lo planes out_hi / out_lo [batch][seq_len][d] (the A operand of
 * a following bevgen_gemm_tc: MaskGit's to_out Linear).  n_cond == seq_len gives dense (bidirectional) attention: MaskGit self-attention
 * runs it with the null key as key 0 and -inf bias entries on the padding keys. */

/* Stage-1 geometric embedding (VQModel.encode with geometric_embedding=True, modules/stage1/vqgan.py:87-109; the default of
 * configs/model/stage_1_cam.yaml): h_nhwc [n_images][hw][d] fp32 += normalize(img_embed(E_inv (I_inv pixel ; 1)) - cam_embed(E_inv[:, 3])) per
 * image (= scene x camera; intrinsics_inv [n_images][3][3], extrinsics_inv [n_images][4][4]) and latent pixel (pixel [hw][3] =
 * (x * image_width, y * image_height, 1)); img_embed_w / cam_embed_w are the [d][4] weights of the two bias-free 1x1 convs.  d <= 1024. */
BEVGEN_API int bevgen_ray_embed_add(float* h_nhwc, const float* intrinsics_inv, const float* extrinsics_inv, const float* pixel, const float* img_embed_w,
                                    const float* cam_embed_w, int n_images, int hw, int d, void* stream);

/* ---------------------------------------------------------------- MaskGit stage-2 variant
 * The bidirectional decoder of modules/stage2/muse_maskgit_pytorch.py runs its Linear layers and the Q.K^T / P.V products on
 * bevgen_gemm_tc, LayerNorm on bevgen_layernorm, the biased softmax on bevgen_attn_softmax; these two produce the operand planes between. */

/* Attention.forward :137-154 (head split, null key / value, cosine-sim normalisation): per (row, head) 64-vector of
 * src[(b*src_batch_rows + r) * src_ld + src_col0 + h*64 ..], optionally x / max(|x|, 1e-12) * scale[0..63] (scale == NULL: plain copy), to bf16
 * hi / lo planes: dst_rows rows per batch element (batch stride dst_batch_rows >= dst_rows rows) with pitch dst_ld elements, heads at columns dst_col0 + 64 h (q | k | v may share one plane).  has_null: destination row 0 = null_vec[h][0..63] (same normalisation), source rows follow
 * from row 1, rows above n_src + 1 are zero (key padding up to the GEMM tile; the caller masks them in the softmax). */
BEVGEN_API int bevgen_mg_head_planes(const float* src, long long src_ld, int src_col0, int n_src, int src_batch_rows, const float* null_vec,
                                     const float* scale, void* out_hi, void* out_lo, int batch, int dst_rows, int dst_batch_rows, long long dst_ld,
                                     int dst_col0, int has_null, int heads, void* stream);
/* FeedForward :72-88 between its Linear layers: u = h[:, f:2f] * gelu(h[:, 0:f]) on rows of pitch h_ld; planes = LayerNorm_f(u) * gamma
 * (eps, biased variance), bf16 hi / lo [rows][f_pad], columns f .. f_pad-1 zero, f_pad <= 3072.  f16f8 != 0: out_hi / out_lo are instead the
 * scaled fp16 plane and the e4m3 pair plane of a following bevgen_linear_f16f8 (f_pad % 64 == 0). */
BEVGEN_API int bevgen_mg_geglu_ln(const float* h, long long h_ld, const float* gamma, void* out_hi, void* out_lo, long long rows, int f, int f_pad,
                                  float eps, int f16f8, void* stream);
/* MaskGit.generate token bookkeeping between the forwards (muse_maskgit_pytorch.py:569-627).
 * mg_sample (:592-603,615-619, top_k / gumbel_sample :41-60): per token row [vocab], keep the top_k logits (ties with the k-th kept),
 * pred = argmax(kept * inv_temperature + gumbel(uniform)), ids[row] = pred where ids[row] == mask_id; scores (may be NULL) receives
 * 1 - softmax(logits)[pred] at masked positions and -1e5 elsewhere.  uniform: [rows][vocab] in (0, 1). */
BEVGEN_API int bevgen_mg_sample(const float* logits, const float* uniform, long long* ids, float* scores, long long rows, int vocab, int top_k,
                                float inv_temperature, long long mask_id, void* stream);
/* mg_remask (:573-582, critic noise :611-613): per camera row [hw], the n_mask largest of scores (+ (uniform - 0.5) * noise_scale when
 * uniform != NULL) are set to mask_id in ids, then positions with init_ids != mask_id (init_ids may be NULL) are restored. */
BEVGEN_API int bevgen_mg_remask(const float* scores, const float* uniform, float noise_scale, long long* ids, const long long* init_ids, long long rows,
                                int hw, int n_mask, long long mask_id, void* stream);

/* ---------------------------------------------------------------- KV-cache autoregressive decode
 * Replaces the per-token full forward of Net2NetTransformer.sample (modules/stage2/cond_transformer_multi_view.py:154-227)
 * with the cached formulation of SURVEY.md §3.4.  All kernels read the step counter s from device memory (one CUDA graph
 * is replayed per token): they process sequence row r = n_cond + s - 1 against keys 0..r and sample decode-order token s.
 * The per-step weight GEMMs are bevgen_gemm_tc launches with BEVGEN_GF_OUT_T (swap-AB, split-K partials [ks][batch][n]). */

/* x = residual + bias + sum_z partials[z]; x_out = x (optional); LayerNorm(x) -> y fp32 (optional) + bf16 planes (optional) */
BEVGEN_API int bevgen_dec_reduce_ln(const float* partials, int ks, long long zstride, const float* bias, const float* residual,
                                    long long residual_row_stride, const float* gamma, const float* beta, float eps, float* x_out, float* y,
                                    void* out_hi, void* out_lo, int rows, int d, void* stream);
/* planes[b][n] = act(bias[n] + sum_z partials[z][b][n]); gelu != 0 -> exact-erf GELU */
BEVGEN_API int bevgen_dec_reduce_act(const float* partials, int ks, long long zstride, const float* bias, int gelu, void* out_hi, void* out_lo,
                                     int rows, int n, void* stream);
/* prefill: rows [0,nrows) of the fused qkv planes [batch][lp][3d] -> K cache [batch][heads][lmax/128][64][128] (blocked, transposed),
 * V cache [batch][heads][lmax][64]; lmax % 128 == 0 */
BEVGEN_API int bevgen_kv_store(const void* qkv_hi, const void* qkv_lo, void* k_cache, void* v_cache, int kv_bf16, int batch, int lp, int nrows,
                               int heads, int d, int lmax, void* stream);
/* one decode row: finish q/k/v (+bias), append k/v, softmax(scale*(q.K + camera_bias[r][:])) V, x1 = y + heads concat.
 * workspace: bevgen_dec_attention_workspace_floats(batch, heads) floats; counters: batch*heads uint32 zero-initialised once.
 * Optional fused LayerNorm of the finished row (ln_gamma != NULL): planes ln_hi/ln_lo[batch][d] = LN(x1); row_counters: batch uint32 zeros. */
BEVGEN_API int bevgen_dec_attention(const float* qkv_partials, int ks, long long zstride, const float* qkv_bias, const float* y,
                                    const float* camera_bias, int bias_ld, void* k_cache, void* v_cache, int kv_bf16, float* x1,
                                    const int* step_ptr, float* workspace, unsigned int* counters, int batch, int n_cond, int heads, int d,
                                    int lmax, float scale, unsigned int* row_counters, const float* ln_gamma, const float* ln_beta, float ln_eps,
                                    void* ln_hi, void* ln_lo, const unsigned char* layout, int layout_block, int layout_ld, void* stream);
/* layout (may be NULL): this layer's per-head block layout uint8 [heads][layout_ld][layout_ld] (density < 1), see bevgen_attn_softmax.
 * k/v cache element type: kv_bf16 = 0 fp32, 1 bf16, 2 fp16. */
BEVGEN_API int bevgen_dec_attention_workspace_floats(int batch, int heads);
/* sampling tail (cond_transformer_multi_view.py:138-142,200-219): logits/T, top-k (ties kept), softmax, multinomial|greedy|forced.
 * forced_tokens[batch][n_img] (decode order, may be NULL): entries >= 0 replace the drawn token (teacher-forced replay, partial decoding of
 * ground-truth cameras :161-165,181-182); negative entries are sampled. */
BEVGEN_API int bevgen_sample_topk(const float* logit_partials, int ks, long long zstride, int vpad, int vocab, float temperature, int top_k,
                                  int greedy, unsigned long long seed, const long long* forced_tokens, const int* forward_shuffle_idx,
                                  long long* cam_idx, long long* tokens_out, float* logits_trace, float* probs_out, const int* step_ptr, int batch,
                                  int n_img, int hw, int ncam, void* stream);
BEVGEN_API int bevgen_dec_advance(int* step_ptr, void* stream);

/* ---------------------------------------------------------------- persistent KV-cache decode kernel
 * ONE launch runs decode steps [step_begin, step_end) of Net2NetTransformer.sample (cond_transformer_multi_view.py:154-227) for up to
 * 16 scenes: per step all transformer blocks (mingpt_sparse.py:240-253 with the single-row attention of sparse_self_attention.py:153-176
 * over the fp16 KV cache), ln_f + head (:385-391), the top-k / softmax / multinomial tail (:200-219) and the embedding of the drawn token
 * (:332-350).  One CTA per SM; weights stream once per step from the packed format below; the layer phases of a step synchronise through
 * generation tags in the data itself (every value that crosses CTAs carries one in its last mantissa bit, the workspace holds every such
 * buffer twice), grid barriers only bracket the head / sampling / embedding of a step.  The
 * caches must have been prefilled (bevgen_kv_store) and token step_begin - 1 must be in cam_idx.  kv caches are fp16. */
typedef struct bevgen_decode_layer {
  const void* w_qkv;                 /* bevgen_pack_decode_linear of [3d][d] (q | k | v rows), columns pre-multiplied by ln1.weight */
  const void* w_1;                   /* of mlp.0.weight [4d][d], columns pre-multiplied by ln2.weight */
  const void* w_2;                   /* of mlp.2.weight [d][4d] with n_quarters = 4 */
  /* lazy LayerNorm: the linears run on the raw residual stream, the epilogue applies rstd_b * (acc - mean_b * c1_n) + c2_n */
  const float* c1_qkv;               /* [3d] sum_k ln1.weight_k * W_nk */
  const float* c2_qkv;               /* [3d] bias_n + sum_k ln1.bias_k * W_nk */
  const float* c2_2;                 /* [d]  mlp.2.bias */
  const float* ln1_g; const float* ln1_b;   /* [d] ln1 itself: Block.forward takes the attention residual from the LayerNorm output */
  const float* c1_1;                 /* [4d] sum_k ln2.weight_k * W1_nk */
  const float* c2_1;                 /* [4d] mlp.0.bias_n + sum_k ln2.bias_k * W1_nk */
  void* k_cache; void* v_cache;      /* this layer's caches, layouts as bevgen_kv_store */
  const unsigned char* layout;       /* optional block layout [heads][layout_ld][layout_ld] */
  float s_qkv, s_1, s_2, pad_;       /* 1 / lo_mul of the three packings */
} bevgen_decode_layer;

typedef struct bevgen_decode_args {
  const bevgen_decode_layer* layers; /* DEVICE array of n_layers entries */
  int n_layers;
  const void* w_head;                /* packed head.weight [vocab][d], columns pre-multiplied by ln_f.weight */
  float s_head;
  const float* c1_head; const float* c2_head;   /* [ceil8(vocab)] lazy ln_f constants (as c1_qkv / c2_qkv, no bias) */
  int batch, d, heads, vocab, n_cond, n_img, lmax, ncam, hw;
  int step_begin, step_end;
  long long* cam_idx;                /* [batch][ncam][hw] */
  const float* x_tok_emb; const float* x_pos_emb; const float* img_embed_w; const float* cam_embed_w;
  const float* intrinsics_inv; const float* extrinsics_inv; const float* pixel;
  const int* forward_shuffle_idx;
  const float* camera_bias; int bias_ld;
  float scale, temperature;
  int top_k, greedy;
  unsigned long long seed;
  const long long* forced_tokens;    /* [batch][n_img] decode order, entries >= 0 replace the draw; may be NULL */
  long long* tokens_out;             /* [batch][n_img] or NULL */
  float* logits_trace;               /* [n_img][batch][vocab] or NULL */
  int layout_block, layout_ld;
  float* workspace;                  /* bevgen_decode_workspace(...) floats (zeroed by the call: all generation tags start at 0) */
  unsigned int* counters;            /* bevgen_decode_workspace(...) uint32 (zeroed by the call) */
  unsigned int* debug;               /* optional pinned HOST buffer of 8 zeroed uint32: a barrier / ring / tag-poll time-out (4 s) leaves (code, CTA, step, layer, phase, ...) here before trapping */
  unsigned long long* profile;       /* optional device buffer [sm_count][32]: ns per phase body / grid barrier, summed over the launch */
} bevgen_decode_args;

BEVGEN_API int bevgen_decode_persistent(const bevgen_decode_args* args, void* stream);
BEVGEN_API int bevgen_decode_workspace(int batch, int d, int heads, int vocab, long long* n_floats, long long* n_counters);
/* Weight packing for the kernel above (a bevgen_pack_* entry point of SURVEY 8b): w [n_rows][ld] fp32 -> 8-row x d-column units in
 * mma.m16n8k16 B-fragment order, fp16 plane + e4m3 plane of (w - fp16(w)) * lo_mul, 3 bytes per weight.  n_quarters > 1 splits the
 * columns into n_quarters ranges of d (MLP2).  Returns the byte count when out == NULL. */
BEVGEN_API long long bevgen_pack_decode_linear(const float* w, int n_rows, int ld, int d, int n_quarters, float lo_mul, void* out, void* stream);


#ifdef __cplusplus
}
#endif
#endif
