// Descriptor of one implicit-GEMM launch (shared between host launcher and kernel).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace bevgen {

constexpr int GEMM_MAX_TAPS = 9;

constexpr int GEMM_FIN_ACT = 1, GEMM_FIN_RESID_LN = 2;

enum GemmFlags : int {
  GF_GELU = 1,        // exact-erf GELU after bias
  GF_OUT_NCHW = 2,    // fp32 output written as [z][col][h][w] (tiny Cout: conv_out)
  GF_B_MN = 4,        // B operand is MN-major in global memory: [k rows][n cols] (e.g. V in P.V)
  GF_CAUSAL_SKIP = 8, // skip output tiles entirely outside the [cond | causal] support (attention scores)
  GF_OUT_T = 32,      // swap-AB decode GEMMs: store D^T, out[z][col][row] fp32, no bias/act/residual (split-K partials)
  GF_OUT_F16F8 = 64,  // out_hi / out_lo are the fp16 plane and the e4m3 pair plane of a following f16f8 GEMM (instead of bf16 hi / lo)
  GF_CAUSAL_KLIMIT = 16, // reduction index = key index: stop at max(ncond, last row of the tile + 1) (P.V)
};

struct GemmParams {
  CUtensorMap tmA[2];   // hi, lo : 4D (c, w, h, n) bf16, box (64, tile_w, tile_h, 1), SWIZZLE_128B
  CUtensorMap tmB[2];   // hi, lo : 2D, K-major: (k, row) box (64, BN); MN-major: (col, krow) box (64, 64)
  int ntaps;
  int tap_dx[GEMM_MAX_TAPS], tap_dy[GEMM_MAX_TAPS], tap_dn[GEMM_MAX_TAPS];
  int a_n_mul, a_n_zstride;   // A image coordinate = z_outer*a_n_mul + z_inner*a_n_zstride + tap_dn[tap]
  int kchunks;          // K / 64 per tap
  int a_c_off, a_c_zstride;         // A channel coordinate = a_c_off + z_inner*a_c_zstride + kc*64
  int b_k_off, b_k_zstride;         // B k coordinate       = b_k_off + z_inner*b_k_zstride + kc*64   (K-major)
  int b_row_zstride, b_row_tapstride;  // B row coordinate  = z_outer*b_row_zstride + tap*b_row_tapstride + n0
  int z_inner, z_outer;
  int tile_w, tile_h, tiles_w, tiles_h;   // M tile = tile_h x tile_w output pixels
  int out_w, out_h;     // valid output extent per z (rows beyond are not stored)
  int n_cols;           // valid output columns (Cout)
  long long out_zo_stride, out_zi_stride;   // element strides of the output per z_outer / z_inner
  int ldc;              // elements between consecutive output rows (pixels)
  const float* bias;    // [n_cols] or null
  const float* residual;  // same indexing as out_f32, or null
  float* out_f32;       // or null
  uint16_t* out_hi;     // bf16 split outputs (same indexing), or null
  uint16_t* out_lo;
  int flags;
  float lo_scale;       // npass == 2: scale of the e4m3 correction accumulator = 1 / (2^13 * S), S = weight scale of ops.pack_f16f8
  int causal_ncond;     // GF_CAUSAL_SKIP: columns < ncond always allowed; else col <= row
  // ---- split-K finalize fused into GF_OUT_T launches (decode): the last CTA to finish a row tile reduces the partials
  int fin_mode;               // 0 none | 1 planes = act(sum + bias) | 2 x = sum + bias + residual, then LayerNorm by the last tile
  int fin_gelu;
  int fin_rows;               // batch rows (columns of D^T)
  const float* fin_bias;      // [n_out]
  const float* fin_resid;     // [rows][n_out]           (mode 2)
  float* fin_x;               // [rows][n_out] fp32      (mode 2)
  float* fin_y;               // LayerNorm output fp32 or null (mode 2)
  uint16_t* fin_hi;           // bf16 planes [rows][n_out] (mode 1: activation; mode 2: LayerNorm output)
  uint16_t* fin_lo;
  const float* fin_gamma;     // (mode 2)
  const float* fin_beta;
  float fin_eps;
  unsigned int* fin_counters; // [m_tiles + 1], zero before the first launch; self-resetting
};

}  // namespace bevgen
