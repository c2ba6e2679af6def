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

// ------------------------------------------------------------------------------------------------
// MaskGit.generate token bookkeeping between the forwards (muse_maskgit_pytorch.py:569-627), two launches per de-masking step.
//
// mg_sample: per token row, `filtered = top-k(logits)`, `pred = argmax(filtered / max(temp, 1e-10) + gumbel(u))` (:592-599, gumbel_sample /
// top_k :41-60 with k = ceil((1 - thres) * vocab)), `ids = where(ids == mask_id, pred, ids)` (:601-603) and, without a token critic, the
// next scores `1 - softmax(logits)[pred]`, -1e5 at positions that were not masked (:615-619).  u is the uniform noise tensor (torch.rand,
// or the reference's own draws in the parity test).  Values tied with the k-th largest logit are all kept.  One CTA of 256 threads per row.
// ------------------------------------------------------------------------------------------------
constexpr int MG_MAXV = 4096;

__global__ void __launch_bounds__(256) mg_sample_kernel(const float* __restrict__ logits, const float* __restrict__ u, long long* __restrict__ ids,
                                                        float* __restrict__ scores, int V, int k, float inv_temp, long long mask_id) {
  __shared__ float lg[MG_MAXV];
  __shared__ int hist[256];
  __shared__ unsigned int sel[2];
  __shared__ float redv[8];
  __shared__ int redi[8];
  const long long row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* lr = logits + row * V;
  float mx = -INFINITY;
  for (int i = tid; i < V; i += 256) { const float v = lr[i]; lg[i] = v; mx = fmaxf(mx, v); }
  __syncthreads();
  // k-th largest logit: 4-pass, 8-bit radix select on the order-preserving integer image of the floats
  float thr = -INFINITY;
  if (k > 0 && k < V) {
    unsigned int prefix = 0u, known = 0u;
    int krem = k;
    for (int pass = 3; pass >= 0; --pass) {
      hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < V; i += 256) {
        const unsigned int b = __float_as_uint(lg[i]);
        const unsigned int key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        if ((key & known) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1);
      }
      __syncthreads();
      if (tid < 32) {                      // lane owns bins 8 lane .. 8 lane + 7; suffix sums run from the top bin down
        int c[8], local = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) { c[e] = hist[tid * 8 + e]; local += c[e]; }
        int incl = local;
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_down_sync(0xffffffffu, incl, o);
          if (tid + o < 32) incl += t;
        }
        const int above = incl - local;
        if (above < krem && krem <= incl) {
          int acc = above;
#pragma unroll
          for (int e = 7; e >= 0; --e) {
            if (acc + c[e] >= krem) { sel[0] = (unsigned)(tid * 8 + e); sel[1] = (unsigned)(krem - acc); break; }
            acc += c[e];
          }
        }
      }
      __syncthreads();
      prefix |= sel[0] << (8 * pass);
      known |= 255u << (8 * pass);
      krem = (int)sel[1];
      __syncthreads();
    }
    thr = __uint_as_float((prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix);
  }
  // arg-max of the kept logits / temp + gumbel noise (lowest index on ties), and the softmax denominator of the raw logits
  const float* ur = u + row * V;
  float best = -INFINITY, se = 0.f;
  int bi = 0x7fffffff;
  for (int i = tid; i < V; i += 256) {
    const float v = lg[i];
    if (v >= thr) {
      const float g = -logf(fmaxf(-logf(fmaxf(ur[i], 1e-20f)), 1e-20f));
      const float t = v * inv_temp + g;
      if (t > best || (t == best && i < bi)) { best = t; bi = i; }
    }
  }
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { redv[w] = best; redi[w] = bi; }
  __syncthreads();
  best = redv[0]; bi = redi[0];
  for (int ww = 1; ww < 8; ++ww)
    if (redv[ww] > best || (redv[ww] == best && redi[ww] < bi)) { best = redv[ww]; bi = redi[ww]; }
  __syncthreads();
  const long long old = ids[row];
  const bool was_mask = old == mask_id;
  if (scores != nullptr) {
    if (lane == 0) redv[w] = mx;
    __syncthreads();
    mx = redv[0];
    for (int ww = 1; ww < 8; ++ww) mx = fmaxf(mx, redv[ww]);
    __syncthreads();
    se = 0.f;
    for (int i = tid; i < V; i += 256) se += expf(lg[i] - mx);
    for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (lane == 0) redv[w] = se;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int ww = 0; ww < 8; ++ww) tot += redv[ww];
      scores[row] = was_mask ? 1.0f - expf(lg[bi] - mx) / tot : -1e5f;
    }
  }
  if (tid == 0 && was_mask) ids[row] = bi;
}

int launch_mg_sample(const float* logits, const float* u, long long* ids, float* scores, long long rows, int V, int k, float inv_temp, long long mask_id,
                     cudaStream_t st) {
  if (rows < 1 || V < 1 || V > MG_MAXV || k < 0) return BEVGEN_ERR_ARG;
  mg_sample_kernel<<<(unsigned)rows, 256, 0, st>>>(logits, u, ids, scores, V, k, inv_temp, mask_id);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

// mg_remask: per camera row of hw tokens, the n_mask positions with the largest scores (+ (u - 0.5) * noise_scale when u is given: the
// critic-noise of :611-613) become mask_id again (:573-579: scores.topk(n_mask).indices scattered into ids), then positions given by
// init_ids (partial decoding: init_ids != mask_id) are restored (:581-582).  Exact selection by rank (ties: lower index first).
constexpr int MG_MAXHW = 2048;

__global__ void __launch_bounds__(256) mg_remask_kernel(const float* __restrict__ scores, const float* __restrict__ u, float noise_scale,
                                                        long long* __restrict__ ids, const long long* __restrict__ init_ids, int hw, int n_mask,
                                                        long long mask_id) {
  __shared__ float sc[MG_MAXHW];
  const long long row = blockIdx.x;
  const int tid = threadIdx.x;
  for (int i = tid; i < hw; i += 256) {
    float v = scores[row * hw + i];
    if (u != nullptr) v += (u[row * hw + i] - 0.5f) * noise_scale;
    sc[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < hw; i += 256) {
    const float v = sc[i];
    int rank = 0;
    for (int j = 0; j < hw; ++j) {
      const float o = sc[j];
      rank += (o > v || (o == v && j < i)) ? 1 : 0;
    }
    long long out = ids[row * hw + i];
    if (rank < n_mask) out = mask_id;
    if (init_ids != nullptr) {
      const long long ini = init_ids[row * hw + i];
      if (ini != mask_id) out = ini;
    }
    ids[row * hw + i] = out;
  }
}

int launch_mg_remask(const float* scores, const float* u, float noise_scale, long long* ids, const long long* init_ids, long long rows, int hw, int n_mask,
                     long long mask_id, cudaStream_t st) {
  if (rows < 1 || hw < 1 || hw > MG_MAXHW || n_mask < 0) return BEVGEN_ERR_ARG;
  mg_remask_kernel<<<(unsigned)rows, 256, 0, st>>>(scores, u, noise_scale, ids, init_ids, hw, n_mask, mask_id);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
