// Stage-2 transformer helper kernels (HBM-bound, coalesced): input-embedding assembly, LayerNorm (+ bf16 split),
// masked/biased attention softmax.  The GEMMs and attention products run in gemm_tc.cu / attn kernels.
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (d % 128 == 0, d <= 1024*4): one warp per row, row kept in registers.
// Writes fp32 y (optional) and bf16 hi/lo planes (optional).  (mingpt_sparse.py:220-221,285)
// ------------------------------------------------------------------------------------------------
template <int NV>  // float4 per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ y, uint16_t* __restrict__ hi,
                                                        uint16_t* __restrict__ lo, long long rows, long long x_row_stride, float eps, int f16f8) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * x_row_stride);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = xr[j * 32 + lane];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / D) + eps);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c4 = j * 32 + lane;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 o;
    o.x = v[j].x * rstd * g.x + b.x; o.y = v[j].y * rstd * g.y + b.y;
    o.z = v[j].z * rstd * g.z + b.z; o.w = v[j].w * rstd * g.w + b.w;
    if (y != nullptr) reinterpret_cast<float4*>(y + row * D)[c4] = o;
    if (hi != nullptr && f16f8) {
      // operand planes of an f16f8 GEMM (gemm_tc npass = 2): fp16 plane + e4m3 pair plane (per 64-column chunk: 64 bytes of 2^13-scaled
      // fp16 remainders, then 64 bytes of values)
      // f16f8 == 2: the fp16 plane holds 2^6 * y (single-accumulator convention of gemm_pair.cu); remainders are against the unscaled value
      const float sc = f16f8 == 2 ? 64.0f : 1.0f, isc = f16f8 == 2 ? 0.015625f : 1.0f;
      const __half2 ha = __floats2half2_rn(o.x * sc, o.y * sc), hb = __floats2half2_rn(o.z * sc, o.w * sc);
      reinterpret_cast<uint2*>(hi + row * D)[c4] = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
      float2 fa = __half22float2(ha), fb = __half22float2(hb);
      fa.x *= isc; fa.y *= isc; fb.x *= isc; fb.y *= isc;
      const uint32_t l8 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((o.x - fa.x) * 8192.0f, (o.y - fa.y) * 8192.0f), __NV_SATFINITE, __NV_E4M3) |
                          ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((o.z - fb.x) * 8192.0f, (o.w - fb.y) * 8192.0f), __NV_SATFINITE, __NV_E4M3) << 16);
      const uint32_t x8 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(o.x, o.y), __NV_SATFINITE, __NV_E4M3) |
                          ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(o.z, o.w), __NV_SATFINITE, __NV_E4M3) << 16);
      const int col = c4 * 4;
      uint8_t* pp = reinterpret_cast<uint8_t*>(lo) + row * (2 * D) + (col >> 6) * 128 + (col & 63);
      *reinterpret_cast<uint32_t*>(pp) = l8;
      *reinterpret_cast<uint32_t*>(pp + 64) = x8;
    } else if (hi != nullptr) {
      __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
      split_bf16(o.x, h0, l0); split_bf16(o.y, h1, l1); split_bf16(o.z, h2, l2); split_bf16(o.w, h3, l3);
      reinterpret_cast<uint2*>(hi + row * D)[c4] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
      if (lo != nullptr) reinterpret_cast<uint2*>(lo + row * D)[c4] = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
    }
  }
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, uint16_t* hi, uint16_t* lo, long long rows, int d,
                     long long x_row_stride, float eps, int f16f8, cudaStream_t st) {
  if (d % 128 != 0 || d > 1024 || rows < 1 || (f16f8 && (!hi || !lo))) return BEVGEN_ERR_ARG;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  switch (d / 128) {
#define LN_CASE(N) case N: layernorm_kernel<N><<<grid, 256, 0, st>>>(x, gamma, beta, y, hi, lo, rows, x_row_stride, eps, f16f8); break;
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
  }
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// Input embeddings (GPT.forward, mingpt_sparse.py:319-373): one CTA (128 threads) per (batch, sequence row).
//   cond row s:  cond_tok_emb[bev[b,s]] + cond_static[s] - sum_cam cam_embed(E_inv[b,cam,:,3])
//   image row :  x_tok_emb[tok] + normalize(img_embed(E_inv (I_inv pixel ; 1)) - cam_embed(E_inv[:,3])) + x_pos_emb[j]
//                with j = forward_shuffle_idx[s - n_cond]  (position embedding is added BEFORE the permutation)
//   pad row   :  x_tok_emb[vocab]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_kernel(const EmbedParams p) {
  __shared__ float red[4];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int s = (p.step_ptr != nullptr ? p.nc + *p.step_ptr - 1 : p.row0) + blockIdx.x;
  float* out = p.out + ((size_t)b * p.nrows + blockIdx.x) * p.d;
  const int tid = threadIdx.x;
  if (s >= p.nc + p.n_img) {  // pad rows
    const float* e = p.x_tok_emb + (size_t)p.vocab * p.d;
    for (int c = tid; c < p.d; c += 128) out[c] = e[c];
    return;
  }
  if (s < p.nc) {
    long long tok = p.bev_idx[(size_t)b * p.nc + s];
    const float* e = p.cond_tok_emb + (size_t)tok * p.d;
    const float* st = p.cond_static + (size_t)s * p.d;
    for (int c = tid; c < p.d; c += 128) {
      float v = e[c] + st[c];
      if (p.bev_embed && p.cam_embed_w != nullptr) {
        float acc = 0.f;
        for (int cam = 0; cam < p.ncam; ++cam) {
          const float* E = p.E_inv + ((size_t)b * p.ncam + cam) * 16;
          const float4 w = __ldg(reinterpret_cast<const float4*>(p.cam_embed_w) + c);
          acc += w.x * E[3] + w.y * E[7] + w.z * E[11] + w.w * E[15];
        }
        v -= acc;
      }
      out[c] = v;
    }
    return;
  }
  const int j = p.fwd[s - p.nc];
  const int cam = j / p.hw, px = j % p.hw;
  long long tok = p.cam_idx[((size_t)b * p.ncam + cam) * p.hw + px];
  if (p.pad_last && j == p.n_img - 1) tok = p.vocab;
  const float* e = p.x_tok_emb + (size_t)tok * p.d;
  const float* pos = p.x_pos_emb + (size_t)j * p.d;
  if (p.img_embed_w == nullptr) {
    for (int c = tid; c < p.d; c += 128) out[c] = e[c] + pos[c];
    return;
  }
  const float* I = p.I_inv + ((size_t)b * p.ncam + cam) * 9;
  const float* E = p.E_inv + ((size_t)b * p.ncam + cam) * 16;
  const float* pix = p.pixel + (size_t)px * 3;
  float cv[4], ray[4];
#pragma unroll
  for (int r = 0; r < 3; ++r) cv[r] = I[r * 3 + 0] * pix[0] + I[r * 3 + 1] * pix[1] + I[r * 3 + 2] * pix[2];
  cv[3] = 1.0f;
#pragma unroll
  for (int r = 0; r < 4; ++r) ray[r] = E[r * 4 + 0] * cv[0] + E[r * 4 + 1] * cv[1] + E[r * 4 + 2] * cv[2] + E[r * 4 + 3] * cv[3];
  // d <= 1024 -> up to 8 values per thread
  float g[8];
  float ss = 0.f;
  int n = 0;
  for (int c = tid; c < p.d; c += 128, ++n) {
    const float4 wi = __ldg(reinterpret_cast<const float4*>(p.img_embed_w) + c);
    const float4 wc = __ldg(reinterpret_cast<const float4*>(p.cam_embed_w) + c);
    const float de = wi.x * ray[0] + wi.y * ray[1] + wi.z * ray[2] + wi.w * ray[3];
    const float ce = wc.x * E[3] + wc.y * E[7] + wc.z * E[11] + wc.w * E[15];
    g[n] = de - ce;
    ss += g[n] * g[n];
  }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  const float inv = 1.0f / (sqrtf(red[0] + red[1] + red[2] + red[3]) + 1e-7f);
  n = 0;
  for (int c = tid; c < p.d; c += 128, ++n) out[c] = (e[c] + g[n] * inv) + pos[c];
}

// ------------------------------------------------------------------------------------------------
// Stage-1 geometric embedding (VQModel.encode with geometric_embedding=True, modules/stage1/vqgan.py:87-109): the encoder output of
// image n = (scene, camera) gets, per latent pixel, the L2-normalised  img_embed(E_inv (I_inv pixel ; 1)) - cam_embed(E_inv[:, 3])
// added in place (h is NHWC fp32).  One CTA (128 threads) per (image, pixel); same arithmetic as the image rows of embed_kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ray_embed_add_kernel(float* __restrict__ h, const float* __restrict__ I_inv, const float* __restrict__ E_inv,
                                                            const float* __restrict__ pixel, const float* __restrict__ img_w,
                                                            const float* __restrict__ cam_w, int hw, int d) {
  __shared__ float red[4];
  const int px = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  float* out = h + ((size_t)n * hw + px) * d;
  const float* I = I_inv + (size_t)n * 9;
  const float* E = E_inv + (size_t)n * 16;
  const float* pix = pixel + (size_t)px * 3;
  float cv[4], ray[4];
#pragma unroll
  for (int r = 0; r < 3; ++r) cv[r] = I[r * 3 + 0] * pix[0] + I[r * 3 + 1] * pix[1] + I[r * 3 + 2] * pix[2];
  cv[3] = 1.0f;
#pragma unroll
  for (int r = 0; r < 4; ++r) ray[r] = E[r * 4 + 0] * cv[0] + E[r * 4 + 1] * cv[1] + E[r * 4 + 2] * cv[2] + E[r * 4 + 3] * cv[3];
  float g[8];
  float ss = 0.f;
  int k = 0;
  for (int c = tid; c < d; c += 128, ++k) {
    const float4 wi = __ldg(reinterpret_cast<const float4*>(img_w) + c);
    const float4 wc = __ldg(reinterpret_cast<const float4*>(cam_w) + c);
    const float de = wi.x * ray[0] + wi.y * ray[1] + wi.z * ray[2] + wi.w * ray[3];
    const float ce = wc.x * E[3] + wc.y * E[7] + wc.z * E[11] + wc.w * E[15];
    g[k] = de - ce;
    ss += g[k] * g[k];
  }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  const float inv = 1.0f / (sqrtf(red[0] + red[1] + red[2] + red[3]) + 1e-7f);
  k = 0;
  for (int c = tid; c < d; c += 128, ++k) out[c] += g[k] * inv;
}

int launch_ray_embed_add(float* h, const float* I_inv, const float* E_inv, const float* pixel, const float* img_w, const float* cam_w, int n_images,
                         int hw, int d, cudaStream_t st) {
  if (d < 1 || d > 1024 || n_images < 1 || n_images > 65535 || hw < 1) return BEVGEN_ERR_ARG;
  ray_embed_add_kernel<<<dim3(hw, n_images), 128, 0, st>>>(h, I_inv, E_inv, pixel, img_w, cam_w, hw, d);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_embed(const EmbedParams& p, cudaStream_t st) {
  if (p.d % 4 != 0 || p.d > 1024 || p.nrows < 1 || p.B < 1 || p.B > 65535) return BEVGEN_ERR_ARG;
  if (launch_k(embed_kernel, dim3(p.nrows, p.B), dim3(128), 0, st, p) != cudaSuccess) return BEVGEN_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// Attention probabilities (sparse_self_attention.py:153-173, dense form):
//   P[z][i][j] = softmax_j( scale * (S[z][i][j] + bias[i][j]) ) over mask[i][j] != 0 ; 0 elsewhere.
// One warp per row, row in registers (L <= 32*MAXJ).  S is never read where masked (may be uninitialised).
// ------------------------------------------------------------------------------------------------
constexpr int ATT_MAXJ = 80;  // L <= 2560

__global__ void __launch_bounds__(256) attn_softmax_kernel(const float* __restrict__ S, const float* __restrict__ bias,
                                                           const uint8_t* __restrict__ mask, uint16_t* __restrict__ hi,
                                                           uint16_t* __restrict__ lo, long long zrows, int L, int Lk, float scale,
                                                           const uint8_t* __restrict__ layout, int H, int blk, int lay_ld) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= zrows) return;
  const int i = (int)(r % L);
  // per-head block layout (sparse_self_attention.py:59-60,153-173, density < 1): key block c / blk of query block i / blk present?
  const uint8_t* lr = layout ? layout + ((size_t)((r / L) % H) * lay_ld + i / blk) * lay_ld : nullptr;
  const float* sr = S + r * Lk;
  const float* br = bias ? bias + (size_t)i * Lk : nullptr;
  const uint8_t* mr = mask + (size_t)i * Lk;
  float v[ATT_MAXJ];
  float m = -INFINITY;
  const int per = (Lk + 31) / 32;
#pragma unroll
  for (int j = 0; j < ATT_MAXJ; ++j) {
    if (j < per) {
      const int c = j * 32 + lane;
      float x = -INFINITY;
      if (c < Lk && mr[c] && (lr == nullptr || lr[c / blk])) x = (sr[c] + (br ? br[c] : 0.f)) * scale;
      v[j] = x;
      m = fmaxf(m, x);
    }
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < ATT_MAXJ; ++j) {
    if (j < per) {
      v[j] = (v[j] == -INFINITY) ? 0.f : expf(v[j] - m);
      sum += v[j];
    }
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < ATT_MAXJ; ++j) {
    if (j < per) {
      const int c = j * 32 + lane;
      if (c < Lk) {
        __nv_bfloat16 h0, l0;
        split_bf16(v[j] * inv, h0, l0);
        hi[r * Lk + c] = __bfloat16_as_ushort(h0);
        if (lo != nullptr) lo[r * Lk + c] = __bfloat16_as_ushort(l0);
      }
    }
  }
}

int launch_attn_softmax(const float* S, const float* bias, const uint8_t* mask, uint16_t* hi, uint16_t* lo, long long zrows, int L, int Lk,
                        float scale, const uint8_t* layout, int H, int blk, int lay_ld, cudaStream_t st) {
  if (Lk > 32 * ATT_MAXJ || Lk < 1 || zrows < 1 || (layout && (H < 1 || blk < 1 || lay_ld * blk < L || lay_ld * blk < Lk))) return BEVGEN_ERR_ARG;
  attn_softmax_kernel<<<(unsigned)((zrows + 7) / 8), 256, 0, st>>>(S, bias, mask, hi, lo, zrows, L, Lk, scale, layout, H, blk, lay_ld);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
