// Direct CUDA-core kernels for the two 3x3 convolutions at the image ends of the VQGAN whose channel count on one side is tiny
// (Encoder.conv_in 3 -> ch, stage1/model.py:355-359; Decoder.conv_out ch -> 3 / 7, :500-504).  As implicit GEMMs they waste the tensor
// core (K = 27 padded to 64, or N = 3 padded to 16) and are bound by operand traffic instead: the im2col plane of conv_in alone is
// 1.6 GB per 96 images, conv_out reads nine shifted boxes of a 6.4 GB operand plane pair.  Here every input element is read once,
// the arithmetic is plain fp32 FMA (exact, no split product), and the GroupNorm statistics / GroupNorm-apply + swish are fused.
#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

// ------------------------------------------------------------------------------------------------
// conv_in: fp32 NCHW [N][3][H][W] -> fp32 NHWC [N][H][W][Cout], Cout = 32 * CPL, bias fused, optional GroupNorm(32) statistics of the
// output (sum / sum of squares per (image, group): a lane owns CPL consecutive channels = exactly one group).
// Warp = one output row segment of 64 pixels at a time (contiguous ranges of segments per warp, so an image change - and with it
// the flush of the statistics - is rare); lane = CPL output channels with their 27 x CPL weights in registers; the 3 x 3 x 66 input
// patch of the segment is staged in shared memory and read as warp-wide broadcasts; 4 pixels are computed per pass so that each
// broadcast feeds up to 3 x 4 FMAs.  Output rows are written as full 128-byte lines (512 B per pixel for Cout = 128).
// ------------------------------------------------------------------------------------------------
constexpr int CI_SEG = 64, CI_PITCH = 68;

template <int CPL>
__global__ void __launch_bounds__(128, 3) conv_in3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ out, double* __restrict__ gn_sums, int N, int H, int W) {
  __shared__ float tile[4][3 * 3 * CI_PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int Cout = 32 * CPL;
  float wr[27][CPL];
#pragma unroll
  for (int k = 0; k < 27; ++k)
#pragma unroll
    for (int j = 0; j < CPL; ++j) wr[k][j] = __ldg(w + (size_t)(lane * CPL + j) * 27 + k);          // OIHW: k = (c*3 + kh)*3 + kw
  float bj[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) bj[j] = bias ? __ldg(bias + lane * CPL + j) : 0.f;
  const int segs_w = (W + CI_SEG - 1) / CI_SEG;
  const long long units = (long long)N * H * segs_w;
  const long long nwarps = (long long)gridDim.x * 4, wid = (long long)blockIdx.x * 4 + warp;
  const long long u0 = units * wid / nwarps, u1 = units * (wid + 1) / nwarps;
  float* tl = tile[warp];
  double ds = 0.0, dq = 0.0;
  int cur_n = -1;
  auto flush = [&](int n) {
    if (gn_sums != nullptr && n >= 0) {
      atomicAdd(gn_sums + (size_t)n * 64 + lane * 2, ds);
      atomicAdd(gn_sums + (size_t)n * 64 + lane * 2 + 1, dq);
    }
    ds = 0.0; dq = 0.0;
  };
  for (long long u = u0; u < u1; ++u) {
    const int sx = (int)(u % segs_w);
    const int h = (int)((u / segs_w) % H);
    const int n = (int)(u / ((long long)segs_w * H));
    if (n != cur_n) { flush(cur_n); cur_n = n; }
    const int x0 = sx * CI_SEG;
    __syncwarp();
    // stage rows h-1 .. h+1, columns x0-1 .. x0+64 of the three input channels (zero padding outside the image)
#pragma unroll
    for (int cr = 0; cr < 9; ++cr) {
      const int c = cr / 3, ih = h + cr % 3 - 1;
      const float* src = x + (((size_t)n * 3 + c) * H + ih) * W;
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const int col = lane + 32 * e;
        if (col < CI_SEG + 2) {
          const int iw = x0 - 1 + col;
          tl[cr * CI_PITCH + col] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(src + iw) : 0.f;
        }
      }
    }
    __syncwarp();
    float s = 0.f, q = 0.f;
#pragma unroll 1
    for (int g = 0; g < CI_SEG / 4; ++g) {
      if (x0 + 4 * g >= W) break;
      float acc[4][CPL];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[p][j] = bj[j];
#pragma unroll
      for (int cr = 0; cr < 9; ++cr) {
        float in[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) in[e] = tl[cr * CI_PITCH + 4 * g + e];          // same address in every lane: broadcast
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[p][j] = fmaf(in[p + kw], wr[cr * 3 + kw][j], acc[p][j]);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int ox = x0 + 4 * g + p;
        if (ox < W) {
          float* op = out + (((size_t)n * H + h) * W + ox) * Cout + lane * CPL;
          if (CPL == 4) *reinterpret_cast<float4*>(op) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
          else if (CPL == 2) *reinterpret_cast<float2*>(op) = make_float2(acc[p][0], acc[p][1]);
          else
#pragma unroll
            for (int j = 0; j < CPL; ++j) op[j] = acc[p][j];
#pragma unroll
          for (int j = 0; j < CPL; ++j) { s += acc[p][j]; q = fmaf(acc[p][j], acc[p][j], q); }
        }
      }
    }
    ds += (double)s;
    dq += (double)q;
  }
  flush(cur_n);
}

int launch_conv_in3(const float* x, const float* w, const float* bias, float* out, double* gn_sums, int N, int H, int W, int Cout, int sm_count,
                    cudaStream_t st) {
  if (N < 1 || H < 1 || W < 1 || !(Cout == 128 || Cout == 64)) return BEVGEN_ERR_ARG;
  const long long units = (long long)N * H * ((W + CI_SEG - 1) / CI_SEG);
  long long ctas = (units + 3) / 4;
  const int grid = (int)(ctas < 3LL * sm_count ? ctas : 3LL * sm_count);
  if (Cout == 128) conv_in3_kernel<4><<<grid, 128, 0, st>>>(x, w, bias, out, gn_sums, N, H, W);
  else conv_in3_kernel<2><<<grid, 128, 0, st>>>(x, w, bias, out, gn_sums, N, H, W);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
