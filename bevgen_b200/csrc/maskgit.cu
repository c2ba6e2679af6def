// Helper kernels of the MaskGit stage-2 variant (reference modules/stage2/muse_maskgit_pytorch.py).  The GEMMs of that
// variant run on gemm_tc, LayerNorm / softmax on transformer.cu; these two kernels are the operand-plane producers in between.
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

// ------------------------------------------------------------------------------------------------
// Attention.forward :137-154: per (row, head) 64-vector, optional cosine-sim normalisation (F.normalize, eps 1e-12) times a learned
// per-channel scale, written as bf16 hi / lo operand planes.  With has_null the destination holds, per batch element, row 0 = the
// head's null key / value (normalised the same way), rows 1 .. n_src = the source rows, rows above = zeros (key padding up to the GEMM
// tile; masked in the softmax).  Source rows of batch b start at row b * src_batch_rows; the destination is dst_rows rows per batch element
// (batch stride dst_batch_rows rows) of pitch dst_ld with the heads at columns dst_col0 + 64 h (so q | k | v can share one fused plane).  One warp per (row, head).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mg_head_planes_kernel(const float* __restrict__ src, long long src_ld, int src_col0, int n_src,
                                                             int src_batch_rows, const float* __restrict__ null_vec,
                                                             const float* __restrict__ scale, uint16_t* __restrict__ hi,
                                                             uint16_t* __restrict__ lo, int dst_rows, int dst_batch_rows, long long dst_ld,
                                                             int dst_col0, int has_null, int H, long long total) {
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= total) return;
  const int lane = threadIdx.x & 31;
  const int h = (int)(w % H);
  const long long drow = w / H;                    // b * dst_rows + r
  const int r = (int)(drow % dst_rows);
  const long long b = drow / dst_rows;
  float2 v = make_float2(0.f, 0.f);
  bool live = true;
  if (has_null && r == 0) {
    v = *reinterpret_cast<const float2*>(null_vec + h * 64 + lane * 2);
  } else if (r - has_null < n_src) {
    v = *reinterpret_cast<const float2*>(src + (b * src_batch_rows + (r - has_null)) * src_ld + src_col0 + h * 64 + lane * 2);
  } else {
    live = false;
  }
  if (live && scale != nullptr) {
    float ss = v.x * v.x + v.y * v.y;
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    const float2 sc = *reinterpret_cast<const float2*>(scale + lane * 2);
    v.x = v.x * inv * sc.x;
    v.y = v.y * inv * sc.y;
  }
  __nv_bfloat16 h0, l0, h1, l1;
  split_bf16(v.x, h0, l0);
  split_bf16(v.y, h1, l1);
  const size_t off = ((size_t)b * dst_batch_rows + r) * dst_ld + dst_col0 + h * 64 + lane * 2;
  *reinterpret_cast<uint32_t*>(hi + off) = pack_bf16(h0, h1);
  if (lo != nullptr) *reinterpret_cast<uint32_t*>(lo + off) = pack_bf16(l0, l1);
}

int launch_mg_head_planes(const float* src, long long src_ld, int src_col0, int n_src, int src_batch_rows, const float* null_vec, const float* scale,
                          uint16_t* hi, uint16_t* lo, int B, int dst_rows, int dst_batch_rows, long long dst_ld, int dst_col0, int has_null, int H,
                          cudaStream_t st) {
  if (B < 1 || H < 1 || dst_rows < n_src + has_null || dst_batch_rows < dst_rows || src_batch_rows < n_src || (has_null && !null_vec) || (src_ld & 1) || (src_col0 & 1) ||
      (dst_ld & 1) || (dst_col0 & 1) || dst_ld < dst_col0 + 64LL * H)
    return BEVGEN_ERR_ARG;
  const long long total = (long long)B * dst_rows * H;
  mg_head_planes_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(src, src_ld, src_col0, n_src, src_batch_rows, null_vec, scale, hi, lo, dst_rows, dst_batch_rows, dst_ld, dst_col0, has_null, H, total);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// FeedForward :72-88 between its two Linear layers: u = gate * gelu(x) with (x | gate) the two halves of the first Linear's output,
// then LayerNorm(f) (gamma only, eps 1e-5, biased variance), written as bf16 hi / lo planes [rows][f_pad] (f_pad = f rounded up to the
// GEMM's 64-wide k chunk, padding = 0).  One CTA per row, the row lives in registers (f <= 256 * 12).
// ------------------------------------------------------------------------------------------------
constexpr int GEGLU_MAX = 12;

// f16f8 != 0: the planes are the scaled f16f8 operand of bevgen_linear_f16f8 instead (fp16(y * 2^6) [rows][f_pad] + e4m3 pair plane
// [rows][2 f_pad bytes]: per 64-column chunk 64 bytes e4m3((y - y16) * 2^13) then 64 bytes e4m3(y)); h_ld = pitch of the input rows.
__global__ void __launch_bounds__(256) mg_geglu_ln_kernel(const float* __restrict__ hin, long long h_ld, const float* __restrict__ gamma,
                                                          uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int f, int f_pad, float eps,
                                                          int f16f8) {
  __shared__ float red[8];
  __shared__ float bc;
  const long long row = blockIdx.x;
  const float* xr = hin + row * h_ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float u[GEGLU_MAX];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < GEGLU_MAX; ++j) {
    const int c = j * 256 + tid;
    u[j] = 0.f;
    if (c < f) {
      u[j] = xr[f + c] * gelu_erf(xr[c]);
      s += u[j];
    }
  }
  auto block_sum = [&](float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) bc = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    __syncthreads();
    return bc;
  };
  const float mean = block_sum(s) / (float)f;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < GEGLU_MAX; ++j) {
    const int c = j * 256 + tid;
    if (c < f) { const float dlt = u[j] - mean; q += dlt * dlt; }
  }
  const float rstd = rsqrtf(block_sum(q) / (float)f + eps);
#pragma unroll
  for (int j = 0; j < GEGLU_MAX; ++j) {
    const int c = j * 256 + tid;
    if (c < f_pad) {
      const float y = c < f ? (u[j] - mean) * rstd * __ldg(gamma + c) : 0.f;
      if (f16f8) {
        const __half hs = __float2half_rn(y * 64.0f);
        reinterpret_cast<__half*>(hi)[row * f_pad + c] = hs;
        uint8_t* pp = reinterpret_cast<uint8_t*>(lo) + row * 2 * f_pad + (c >> 6) * 128 + (c & 63);
        pp[0] = (uint8_t)__nv_cvt_float_to_fp8(fmaf(__half2float(hs), -128.0f, y * 8192.0f), __NV_SATFINITE, __NV_E4M3);
        pp[64] = (uint8_t)__nv_cvt_float_to_fp8(y, __NV_SATFINITE, __NV_E4M3);
        continue;
      }
      __nv_bfloat16 h0, l0;
      split_bf16(y, h0, l0);
      hi[row * f_pad + c] = __bfloat16_as_ushort(h0);
      if (lo != nullptr) lo[row * f_pad + c] = __bfloat16_as_ushort(l0);
    }
  }
}

int launch_mg_geglu_ln(const float* hin, long long h_ld, const float* gamma, uint16_t* hi, uint16_t* lo, long long rows, int f, int f_pad, float eps,
                       int f16f8, cudaStream_t st) {
  if (rows < 1 || f < 1 || f_pad < f || f_pad > 256 * GEGLU_MAX || h_ld < 2LL * f || (f16f8 && (!lo || (f_pad & 63)))) return BEVGEN_ERR_ARG;
  mg_geglu_ln_kernel<<<(unsigned)rows, 256, 0, st>>>(hin, h_ld, gamma, hi, lo, f, f_pad, eps, f16f8);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
