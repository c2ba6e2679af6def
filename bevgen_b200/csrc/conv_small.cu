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

// ------------------------------------------------------------------------------------------------
// conv_out: y = conv3x3(swish(GroupNorm(x))) for Cout = 3: fp32 NHWC [N][H][W][C] (C = 32 * CPL) -> fp32 NCHW [N][3][H][W].
// CTA (4 warps) = 4 output rows x 16 pixels.  Staging: the (4+2) x (16+2) x C halo is read once (thread = (pixel, channel quad):
// full 512-byte pixel rows per warp), GroupNorm affine + swish applied, stored as [pixel][C] in shared memory (zero padding after the
// transform).  Compute: warp = output row, lane = CPL input channels with their 9 x CPL x 3 weights in registers; 4 pixels per pass
// (a 3 x 6 patch of 16-byte conflict-free shared loads feeds 4 x 27 x CPL FMAs), the 12 per-lane partial sums are reduced over the
// 32 channel lanes with a transposing butterfly (16 shuffles) and written straight to the NCHW image.
// ------------------------------------------------------------------------------------------------
constexpr int CO_TW = 16, CO_TH = 4, CO_HW = CO_TW + 2, CO_HH = CO_TH + 2;

template <int CPL>
__global__ void __launch_bounds__(128, 3) conv_out3_kernel(const float* __restrict__ x, const float* __restrict__ affine, int swish,
                                                           const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                                           int N, int H, int W) {
  constexpr int C = 32 * CPL;
  extern __shared__ __align__(16) float halo[];            // [CO_HH * CO_HW pixels][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wr[27][CPL];                                       // [co*9 + tap][own channel]
#pragma unroll
  for (int co = 0; co < 3; ++co)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int j = 0; j < CPL; ++j) wr[co * 9 + tap][j] = __ldg(w + ((size_t)co * C + lane * CPL + j) * 9 + tap);      // OIHW
  const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f, b2 = bias ? __ldg(bias + 2) : 0.f;
  const int tiles_w = (W + CO_TW - 1) / CO_TW, tiles_h = (H + CO_TH - 1) / CO_TH;
  const long long tiles = (long long)N * tiles_w * tiles_h;
  int aff_n = -1;
  float sc[CPL], sf[CPL];
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int tw = (int)(t % tiles_w), th = (int)((t / tiles_w) % tiles_h), n = (int)(t / ((long long)tiles_w * tiles_h));
    if (n != aff_n) {                                      // this lane's GroupNorm (scale, shift), once per image
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        sc[j] = affine ? __ldg(affine + ((size_t)n * C + lane * CPL + j) * 2) : 1.f;
        sf[j] = affine ? __ldg(affine + ((size_t)n * C + lane * CPL + j) * 2 + 1) : 0.f;
      }
      aff_n = n;
    }
    __syncthreads();                                       // previous tile fully consumed
    // ---- staging: warp handles halo pixels warp, warp + 4, ...; lane = its CPL channels (the same channels it owns in the compute phase).
    // Loads are issued nine pixels at a time before any of them is consumed (27 pixels per warp = 3 batches): memory-level parallelism.
    constexpr int SB = 9;
    static_assert((CO_HH * CO_HW) % (4 * SB) == 0, "halo pixels must split into whole batches");
#pragma unroll 1
    for (int b0 = 0; b0 < CO_HH * CO_HW / 4; b0 += SB) {
      float v[SB][CPL];
      bool ok[SB];
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        const int hp = warp + 4 * (b0 + i);
        const int gh = th * CO_TH - 1 + hp / CO_HW, gw = tw * CO_TW - 1 + hp % CO_HW;
        ok[i] = gh >= 0 && gh < H && gw >= 0 && gw < W;
        const float* src = x + (((size_t)n * H + (ok[i] ? gh : 0)) * W + (ok[i] ? gw : 0)) * C + lane * CPL;
        if (CPL == 4) { const float4 q = *reinterpret_cast<const float4*>(src); v[i][0] = q.x; v[i][1] = q.y; v[i][2] = q.z; v[i][3] = q.w; }
        else { const float2 q = *reinterpret_cast<const float2*>(src); v[i][0] = q.x; v[i][1] = q.y; }
      }
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        const int hp = warp + 4 * (b0 + i);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          float a = fmaf(v[i][j], sc[j], sf[j]);
          if (swish) {
            float ex, rc;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(a * -1.4426950408889634f));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
            a *= rc;
          }
          v[i][j] = ok[i] ? a : 0.f;                        // zero padding applies after the transform
        }
        float* dst = halo + (size_t)hp * C + lane * CPL;
        if (CPL == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        else *reinterpret_cast<float2*>(dst) = make_float2(v[i][0], v[i][1]);
      }
    }
    __syncthreads();
    // ---- compute: warp = output row th*4 + warp
    const int oh = th * CO_TH + warp;
#pragma unroll 1
    for (int g = 0; g < CO_TW / 4; ++g) {
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;             // [co * 4 + px], 12 used
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        float in[6][CPL];
#pragma unroll
        for (int e = 0; e < 6; ++e) {
          const float* src = halo + (size_t)((warp + kh) * CO_HW + 4 * g + e) * C + lane * CPL;
          if (CPL == 4) { const float4 q = *reinterpret_cast<const float4*>(src); in[e][0] = q.x; in[e][1] = q.y; in[e][2] = q.z; in[e][3] = q.w; }
          else { const float2 q = *reinterpret_cast<const float2*>(src); in[e][0] = q.x; in[e][1] = q.y; }
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int co = 0; co < 3; ++co)
#pragma unroll
            for (int px = 0; px < 4; ++px)
#pragma unroll
              for (int j = 0; j < CPL; ++j) acc[co * 4 + px] = fmaf(in[px + kw][j], wr[co * 9 + kh * 3 + kw][j], acc[co * 4 + px]);
      }
      // transposing butterfly over the 16 slots: lane l ends up with slot (l >> 1) summed over all 32 channel lanes
      int nlive = 16;
#pragma unroll
      for (int off = 16; off >= 2; off >>= 1) {
        const int half = nlive / 2;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < half) {
            const float send = up ? acc[i] : acc[i + half];
            const float keep = up ? acc[i + half] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        nlive = half;
      }
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
      const int slot = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
      const int co = slot >> 2, px = slot & 3, ox = tw * CO_TW + 4 * g + px;
      if (!(lane & 1) && co < 3 && oh < H && ox < W)
        out[(((size_t)n * 3 + co) * H + oh) * W + ox] = acc[0] + (co == 0 ? b0 : (co == 1 ? b1 : b2));
    }
  }
}

int launch_conv_out3(const float* x, const float* affine, int swish, const float* w, const float* bias, float* out, int N, int H, int W, int C,
                     int sm_count, cudaStream_t st) {
  if (N < 1 || H < 1 || W < 1 || !(C == 128 || C == 64)) return BEVGEN_ERR_ARG;
  const long long tiles = (long long)N * ((W + CO_TW - 1) / CO_TW) * ((H + CO_TH - 1) / CO_TH);
  const int grid = (int)(tiles < 3LL * sm_count ? tiles : 3LL * sm_count);
  const int smem = CO_HH * CO_HW * C * 4;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv_out3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, CO_HH * CO_HW * 128 * 4) != cudaSuccess ||
        cudaFuncSetAttribute(conv_out3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CO_HH * CO_HW * 64 * 4) != cudaSuccess)
      return BEVGEN_ERR_CUDA;
    configured = true;
  }
  if (C == 128) conv_out3_kernel<4><<<grid, 128, smem, st>>>(x, affine, swish, w, bias, out, N, H, W);
  else conv_out3_kernel<2><<<grid, 128, smem, st>>>(x, affine, swish, w, bias, out, N, H, W);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
