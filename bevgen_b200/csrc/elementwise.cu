// HBM-bound helper kernels around the tensor-core GEMM: GroupNorm statistics, the "prep" pass that turns an
// fp32 NHWC activation into the bf16 (hi, lo) operand planes the GEMM consumes (fusing GroupNorm-apply, swish,
// nearest 2x upsampling or the stride-2 space-to-depth split), conv_in im2col, LayerNorm, layout transposes,
// row softmax and row gathers.  All are coalesced / 128-bit vectorised grid-stride kernels.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

// ------------------------------------------------------------------------------------------------
// GroupNorm(32 groups) statistics over fp32 NHWC [n][pixels][C] -> double sums[n][32][2] (sum, sumsq).
// blockDim = 256; threads split as (C/4 channel quads) x (256/(C/4) pixels); grid = (chunks, n).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, int pixels, int C,
                                                       int pix_per_block) {
  extern __shared__ double sm[];  // [C][2]
  const int n = blockIdx.y;
  const int quads = C >> 2;
  const int tq = threadIdx.x % quads, tp = threadIdx.x / quads;
  const int pstep = blockDim.x / quads;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, pixels);
  const float4* base = reinterpret_cast<const float4*>(x + (size_t)n * pixels * C);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  double ds[4] = {0, 0, 0, 0}, dss[4] = {0, 0, 0, 0};
  int cnt = 0;
  for (int p = p0 + tp; p < p1; p += pstep) {
    float4 v = __ldg(base + (size_t)p * quads + tq);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    ss[0] += v.x * v.x; ss[1] += v.y * v.y; ss[2] += v.z * v.z; ss[3] += v.w * v.w;
    if (++cnt == 32) {  // bound fp32 accumulation error: flush to double every 32 terms
#pragma unroll
      for (int j = 0; j < 4; ++j) { ds[j] += s[j]; dss[j] += ss[j]; s[j] = 0; ss[j] = 0; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    atomicAdd(&sm[(tq * 4 + j) * 2 + 0], ds[j] + (double)s[j]);
    atomicAdd(&sm[(tq * 4 + j) * 2 + 1], dss[j] + (double)ss[j]);
  }
  __syncthreads();
  const int cpg = C / 32;
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    double a = 0;
    for (int c = 0; c < cpg; ++c) a += sm[(g * cpg + c) * 2 + which];
    atomicAdd(&sums[((size_t)n * 32 + g) * 2 + which], a);
  }
}

// sums -> (mean, rstd) per (n, group); count = pixels * channels-per-group
__global__ void gn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ mean_rstd, int n_groups_total, double inv_cnt,
                                   double eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_groups_total) return;
  const double mean = sums[2 * i] * inv_cnt;
  double var = sums[2 * i + 1] * inv_cnt - mean * mean;
  var = var < 0 ? 0 : var;
  mean_rstd[2 * i] = (float)mean;
  mean_rstd[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

// ------------------------------------------------------------------------------------------------
// prep: fp32 NHWC -> bf16 hi/lo planes, optional GroupNorm-apply (+swish), optional spatial remap.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_kernel(const PrepParams p) {
  const int oct = p.C >> 3;  // 8 channels per thread
  int OH = p.H, OW = p.W, ON = p.N;
  if (p.mode == PREP_UP2) { OH = 2 * p.H; OW = 2 * p.W; }
  if (p.mode == PREP_S2D) { OH = p.H / 2; OW = p.W / 2; ON = p.N * 4; }
  const size_t total = (size_t)ON * OH * OW * oct;
  const int cpg = p.C / 32;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % oct);
    size_t r = i / oct;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int on = (int)(r / OH);
    int n = on, h = oh, w = ow;
    if (p.mode == PREP_UP2) { h = oh >> 1; w = ow >> 1; }
    if (p.mode == PREP_S2D) { n = on >> 2; h = 2 * oh + ((on >> 1) & 1); w = 2 * ow + (on & 1); }
    const float4* src = reinterpret_cast<const float4*>(p.x + (((size_t)n * p.H + h) * p.W + w) * p.C + c8 * 8);
    float4 a = __ldg(src), b = __ldg(src + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (p.mean_rstd != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c8 * 8 + j;
        const float2 mr = __ldg(reinterpret_cast<const float2*>(p.mean_rstd) + (size_t)n * 32 + c / cpg);
        v[j] = (v[j] - mr.x) * mr.y * __ldg(p.gamma + c) + __ldg(p.beta + c);
      }
    }
    if (p.swish) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] / (1.0f + expf(-v[j]));
    }
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * e], h0, l0);
      split_bf16(v[2 * e + 1], h1, l1);
      hh[e] = pack_bf16(h0, h1);
      ll[e] = pack_bf16(l0, l1);
    }
    reinterpret_cast<uint4*>(p.hi)[i] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    if (p.lo != nullptr) reinterpret_cast<uint4*>(p.lo)[i] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// conv_in im2col: fp32 NCHW [N][Cin][H][W] (Cin*9 <= 64) -> bf16 hi/lo [N][H][W][64], k = (kh*3+kw)*Cin + c
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                        int N, int Cin, int H, int W) {
  const size_t total = (size_t)N * H * W * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k8 = (int)(i & 7);
    size_t r = i >> 3;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k = k8 * 8 + 2 * e + u;
        float val = 0.f;
        if (k < 9 * Cin) {
          const int tap = k / Cin, c = k % Cin;
          const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = __ldg(x + (((size_t)n * Cin + c) * H + ih) * W + iw);
        }
        v2[u] = val;
      }
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v2[0], h0, l0);
      split_bf16(v2[1], h1, l1);
      hh[e] = pack_bf16(h0, h1);
      ll[e] = pack_bf16(l0, l1);
    }
    reinterpret_cast<uint4*>(hi)[i] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    if (lo != nullptr) reinterpret_cast<uint4*>(lo)[i] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 layout transposes [N][C][P] <-> [N][P][C] via a 32x32 smem tile (both sides coalesced)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc) {
  // src [n][R][Cc] -> dst [n][Cc][R]
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const float* s = src + (size_t)n * R * Cc;
  float* d = dst + (size_t)n * R * Cc;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < Cc) ? s[(size_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, r = r0 + tx;
    if (r < R && c < Cc) d[(size_t)c * R + r] = tile[tx][j];
  }
}

// ------------------------------------------------------------------------------------------------
// Row softmax of scale*x over fp32 [rows][cols] (cols <= 1024) -> bf16 hi/lo probabilities (VQGAN AttnBlock)
// one warp per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                           long long rows, int cols, int out_ld, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * cols;
  float v[32];
  float m = -INFINITY;
  const int per = (cols + 31) / 32;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < per) {
      int c = j * 32 + lane;
      v[j] = (c < cols) ? sr[c] * scale : -INFINITY;
      m = fmaxf(m, v[j]);
    }
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < per) {
      v[j] = expf(v[j] - m);
      sum += v[j];
    }
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < per) {
      int c = j * 32 + lane;
      if (c < cols) {
        __nv_bfloat16 h0, l0;
        split_bf16(v[j] * inv, h0, l0);
        hi[row * out_ld + c] = __bfloat16_as_ushort(h0);
        if (lo != nullptr) lo[row * out_ld + c] = __bfloat16_as_ushort(l0);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather: out[r][:] = table[idx[r]][:]  (codebook lookup, quantize.py:314-329), fp32, D % 4 == 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx,
                                                          float* __restrict__ out, long long rows, int D, int n_table) {
  const int q = D >> 2;
  const size_t total = (size_t)rows * q;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const long long r = i / q;
    const int c = (int)(i % q);
    long long id = idx[r];
    id = id < 0 ? 0 : (id >= n_table ? n_table - 1 : id);
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(table + (size_t)id * D) + c);
  }
}

// ------------------------------------------------------------------------------------------------
// denormalize: x*std+mean per channel, clamp [0,1] on fp32 NCHW (bev_utils/util.py:97-118)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) denorm_kernel(const float* __restrict__ x, float* __restrict__ out, size_t total, int C, int P,
                                                     float m0, float m1, float m2, float s0, float s1, float s2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / P) % C);
    const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), s = c == 0 ? s0 : (c == 1 ? s1 : s2);
    out[i] = fminf(fmaxf(x[i] * s + m, 0.f), 1.f);
  }
}

// ---------------------------------------------------------------- launchers
static inline int grid_for(size_t total, int block, int sm_count) {
  size_t g = (total + block - 1) / block;
  size_t cap = (size_t)sm_count * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

int launch_gn_stats(const float* x, double* sums, float* mean_rstd, int N, int pixels, int C, float eps, cudaStream_t st) {
  if (C % 32 != 0 || C > 1024 || (256 % (C / 4)) != 0) return BEVGEN_ERR_ARG;
  if (cudaMemsetAsync(sums, 0, (size_t)N * 64 * sizeof(double), st) != cudaSuccess) return BEVGEN_ERR_CUDA;
  const int pstep = 256 / (C / 4);
  int pix_per_block = pstep * 64;
  int chunks = (pixels + pix_per_block - 1) / pix_per_block;
  gn_stats_kernel<<<dim3(chunks, N), 256, 2 * C * sizeof(double), st>>>(x, sums, pixels, C, pix_per_block);
  gn_finalize_kernel<<<(N * 32 + 127) / 128, 128, 0, st>>>(sums, mean_rstd, N * 32, 1.0 / ((double)pixels * (C / 32)), (double)eps);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// sums -> per-(image, channel) affine of the fused GroupNorm: y = x * scale + shift, scale = rstd*gamma, shift = beta - mean*rstd*gamma
__global__ void gn_affine_kernel(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ affine, int N, int C, double inv_cnt, double eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C, g = c / (C / 32);
  const double mean = sums[((size_t)n * 32 + g) * 2] * inv_cnt;
  double var = sums[((size_t)n * 32 + g) * 2 + 1] * inv_cnt - mean * mean;
  var = var < 0 ? 0 : var;
  const double rstd = 1.0 / sqrt(var + eps);
  const double sc = rstd * (double)gamma[c];
  affine[2 * i] = (float)sc;
  affine[2 * i + 1] = (float)((double)beta[c] - mean * sc);
}

int launch_gn_affine(const double* sums, const float* gamma, const float* beta, float* affine, int N, int pixels, int C, float eps, cudaStream_t st) {
  if (C % 32 != 0) return BEVGEN_ERR_ARG;
  gn_affine_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(sums, gamma, beta, affine, N, C, 1.0 / ((double)pixels * (C / 32)), (double)eps);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_gn_finalize(const double* sums, float* mean_rstd, int N, int pixels, int C, float eps, cudaStream_t st) {
  gn_finalize_kernel<<<(N * 32 + 127) / 128, 128, 0, st>>>(sums, mean_rstd, N * 32, 1.0 / ((double)pixels * (C / 32)), (double)eps);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_prep(const PrepParams& p, int sm_count, cudaStream_t st) {
  if (p.C % 8 != 0) return BEVGEN_ERR_ARG;
  if (p.mean_rstd != nullptr && p.C % 32 != 0) return BEVGEN_ERR_ARG;
  if (p.mode == PREP_S2D && ((p.H | p.W) & 1)) return BEVGEN_ERR_ARG;
  size_t out_pix = (size_t)p.N * p.H * p.W * (p.mode == PREP_UP2 ? 4 : 1);
  prep_kernel<<<grid_for(out_pix * (p.C / 8), 256, sm_count), 256, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_im2col3x3(const float* x, uint16_t* hi, uint16_t* lo, int N, int Cin, int H, int W, int sm_count, cudaStream_t st) {
  if (Cin * 9 > 64 || Cin < 1) return BEVGEN_ERR_ARG;
  im2col3x3_kernel<<<grid_for((size_t)N * H * W * 8, 256, sm_count), 256, 0, st>>>(x, hi, lo, N, Cin, H, W);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_transpose(const float* src, float* dst, int N, int R, int Cc, cudaStream_t st) {
  if (N > 65535) return BEVGEN_ERR_ARG;
  transpose_kernel<<<dim3((Cc + 31) / 32, (R + 31) / 32, N), 256, 0, st>>>(src, dst, R, Cc);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_softmax_rows(const float* s, uint16_t* hi, uint16_t* lo, long long rows, int cols, int out_ld, float scale, cudaStream_t st) {
  if (cols > 1024 || cols < 1) return BEVGEN_ERR_ARG;
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(s, hi, lo, rows, cols, out_ld, scale);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_gather_rows(const float* table, const long long* idx, float* out, long long rows, int D, int n_table, int sm_count,
                       cudaStream_t st) {
  if (D % 4 != 0) return BEVGEN_ERR_ARG;
  gather_rows_kernel<<<grid_for((size_t)rows * (D / 4), 256, sm_count), 256, 0, st>>>(table, idx, out, rows, D, n_table);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// fp32 NCHW images in [0,1] -> uint8 NHWC (round to nearest), the layout image encoders want; a quarter of the D2H bytes of the fp32 tensor.
__global__ void to_uint8_hwc_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, size_t total, int C, int P) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / ((size_t)C * P), r = i % ((size_t)C * P);
    const int px = (int)(r / C), c = (int)(r % C);                      // output order: pixel-major, channel fastest
    const float v = __ldg(x + (n * C + c) * P + px);
    out[i] = (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * 255.0f);
  }
}

int launch_to_uint8_hwc(const float* x, uint8_t* out, int N, int C, int P, int sm_count, cudaStream_t st) {
  if (N < 1 || C < 1 || P < 1) return BEVGEN_ERR_ARG;
  const size_t total = (size_t)N * C * P;
  to_uint8_hwc_kernel<<<grid_for(total, 256, sm_count), 256, 0, st>>>(x, out, total, C, P);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_denorm(const float* x, float* out, int N, int C, int P, const float* mean, const float* std_, int sm_count, cudaStream_t st) {
  if (C != 3) return BEVGEN_ERR_ARG;
  size_t total = (size_t)N * C * P;
  denorm_kernel<<<grid_for(total, 256, sm_count), 256, 0, st>>>(x, out, total, C, P, mean[0], mean[1], mean[2], std_[0], std_[1], std_[2]);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ weight packing (load time)
// max |x| over n floats -> *out (a float that must hold 0 before the launch; values are non-negative, so their bit patterns order like ints)
__global__ void absmax_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
int launch_absmax(const float* x, long long n, float* out, int sm_count, cudaStream_t st) {
  absmax_kernel<<<sm_count * 4, 256, 0, st>>>(x, n, out);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// fp32 -> bf16 hi (+ bf16 lo = bf16(x - hi)) planes (the operand format of the bf16x3 split product)
__global__ void split_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo != nullptr) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
int launch_split_bf16(const float* x, long long n, void* hi, void* lo, int sm_count, cudaStream_t st) {
  split_bf16_kernel<<<sm_count * 8, 256, 0, st>>>(x, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// fp32 weight rows [rows][cin] -> the f16f8 weight operand: w16 = fp16(w * w16_mul) and, per `chunk`-element piece of a row, `chunk` bytes
// e4m3(w * s) followed by `chunk` bytes e4m3((w - w16 / w16_mul) * s * 2^13) (saturating, round to nearest even).  One thread per element.
__global__ void pack_f16f8_kernel(const float* __restrict__ w, long long n, int cin, int chunk, float s, float w16_mul, __half* __restrict__ w16,
                                  uint8_t* __restrict__ pair) {
  const float lo_mul = s * 8192.0f, inv16 = 1.0f / w16_mul;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = w[i];
    const __half h = __float2half_rn(v * w16_mul);
    w16[i] = h;
    const long long row = i / cin;
    const int c = (int)(i - row * cin), piece = c / chunk, e = c - piece * chunk;
    uint8_t* dst = pair + row * 2 * cin + (long long)piece * 2 * chunk + e;
    dst[0] = (uint8_t)__nv_cvt_float_to_fp8(v * s, __NV_SATFINITE, __NV_E4M3);
    dst[chunk] = (uint8_t)__nv_cvt_float_to_fp8((v - __half2float(h) * inv16) * lo_mul, __NV_SATFINITE, __NV_E4M3);
  }
}
int launch_pack_f16f8(const float* w, long long rows, int cin, int chunk, float s, float w16_mul, void* w16, void* pair, int sm_count, cudaStream_t st) {
  if (rows < 1 || cin < chunk || cin % chunk != 0 || !(chunk == 32 || chunk == 64) || !(s > 0.f) || !(w16_mul > 0.f)) return BEVGEN_ERR_ARG;
  pack_f16f8_kernel<<<sm_count * 8, 256, 0, st>>>(w, rows * cin, cin, chunk, s, w16_mul, (__half*)w16, (uint8_t*)pair);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
