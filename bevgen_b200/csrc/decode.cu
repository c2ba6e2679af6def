// KV-cache autoregressive decode kernels (the reference has no KV cache: Net2NetTransformer.sample re-runs the full
// 1792-token forward per token, modules/stage2/cond_transformer_multi_view.py:172-219; SURVEY.md §3.4 derives the
// equivalent cached formulation from the closed-form mask).  Every kernel reads the current step from device memory so
// one CUDA graph can be replayed for all 1536 tokens.  Step s samples decode-order token s; it processes sequence row
// r = n_cond + s - 1 (the row holding token s-1, or the last cond row for s = 0) against keys 0..r.
// The weight GEMMs of a step are swap-AB tcgen05 launches (gemm_tc.cu, GF_OUT_T) that leave split-K partials
// [ks][batch][features]; the kernels here fold "sum partials + bias (+ activation / residual / LayerNorm)" into one pass.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max_256(float v, float* red) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
  return t;
}

// ------------------------------------------------------------------------------------------------
// x = residual + bias + sum_z partials[z] ; (optional) x_out = x ; y = LayerNorm(x) -> fp32 + bf16 planes.
// One CTA of 256 threads per batch row; d <= 1024, d % 4 == 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_reduce_ln_kernel(const float* __restrict__ partials, int ks, long long zstride,
                                                            const float* __restrict__ bias, const float* __restrict__ residual,
                                                            long long residual_row_stride, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ x_out,
                                                            float* __restrict__ y, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int d) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, c4 = threadIdx.x;
  const bool act = c4 < (d >> 2);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) {
    if (residual != nullptr) v = reinterpret_cast<const float4*>(residual + (size_t)row * residual_row_stride)[c4];
    if (bias != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    int z = 0;
    for (; z + 4 <= ks; z += 4) {       // 4 independent loads in flight
      const float4 t0 = reinterpret_cast<const float4*>(partials + (z + 0) * zstride + (size_t)row * d)[c4];
      const float4 t1 = reinterpret_cast<const float4*>(partials + (z + 1) * zstride + (size_t)row * d)[c4];
      const float4 t2 = reinterpret_cast<const float4*>(partials + (z + 2) * zstride + (size_t)row * d)[c4];
      const float4 t3 = reinterpret_cast<const float4*>(partials + (z + 3) * zstride + (size_t)row * d)[c4];
      v.x += (t0.x + t1.x) + (t2.x + t3.x); v.y += (t0.y + t1.y) + (t2.y + t3.y);
      v.z += (t0.z + t1.z) + (t2.z + t3.z); v.w += (t0.w + t1.w) + (t2.w + t3.w);
    }
    for (; z < ks; ++z) {
      const float4 t = reinterpret_cast<const float4*>(partials + z * zstride + (size_t)row * d)[c4];
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (x_out != nullptr) reinterpret_cast<float4*>(x_out + (size_t)row * d)[c4] = v;
  }
  const float mean = block_sum_256(act ? (v.x + v.y) + (v.z + v.w) : 0.f, red) / d;
  v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
  const float var = block_sum_256(act ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f, red) / d;
  const float rstd = rsqrtf(var + eps);
  if (!act) return;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  float4 o;
  o.x = v.x * rstd * g.x + b.x; o.y = v.y * rstd * g.y + b.y; o.z = v.z * rstd * g.z + b.z; o.w = v.w * rstd * g.w + b.w;
  if (y != nullptr) reinterpret_cast<float4*>(y + (size_t)row * d)[c4] = o;
  if (hi != nullptr) {
    __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
    split_bf16(o.x, h0, l0); split_bf16(o.y, h1, l1); split_bf16(o.z, h2, l2); split_bf16(o.w, h3, l3);
    reinterpret_cast<uint2*>(hi + (size_t)row * d)[c4] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
    if (lo != nullptr) reinterpret_cast<uint2*>(lo + (size_t)row * d)[c4] = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
  }
}

// planes[b][n] = act(bias[n] + sum_z partials[z][b][n])   (MLP hidden: exact-erf GELU)
__global__ void __launch_bounds__(256) dec_reduce_act_kernel(const float* __restrict__ partials, int ks, long long zstride,
                                                             const float* __restrict__ bias, uint16_t* __restrict__ hi,
                                                             uint16_t* __restrict__ lo, int rows, int n, int gelu) {
  pdl_launch_dependents();
  pdl_wait();
  const int q = n >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * q) return;
  const int c4 = i % q;
  float4 v = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  for (int z = 0; z < ks; ++z) {
    const float4 t = reinterpret_cast<const float4*>(partials + z * zstride)[i];
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  if (gelu) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
  __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
  split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
  reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
  if (lo != nullptr) reinterpret_cast<uint2*>(lo)[i] = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
}

// ------------------------------------------------------------------------------------------------
// KV cache.  K is stored transposed per (batch, head): [64][Lmax] so that the score pass (thread = key) is coalesced;
// V is stored [Lmax][64] so that the P.V pass (thread = channel) is coalesced.  KVT = float (exact) or bf16 (fast mode).
// ------------------------------------------------------------------------------------------------
// K cache index: blocked by 128 keys so that the (64 x 128) slab a decode CTA needs is one contiguous 32 KB (fp32) block:
//   K[b][h][j / 128][c][j % 128]          (Lmax must be a multiple of 128)
__device__ __forceinline__ size_t k_index(size_t bh, int c, int j, int Lmax) {
  return ((bh * (size_t)(Lmax >> 7) + (size_t)(j >> 7)) * 64 + (size_t)c) * 128 + (size_t)(j & 127);
}

template <typename T> __device__ __forceinline__ float kv_load(const T* p);
template <> __device__ __forceinline__ float kv_load<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float kv_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float kv_load<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void kv_store(T* p, float v);
template <> __device__ __forceinline__ void kv_store<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void kv_store<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void kv_store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// prefill: copy rows [0, nrows) of the fused qkv planes [B][Lp][3d] (value = hi + lo) into the caches
template <typename KVT>
__global__ void __launch_bounds__(256) kv_store_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, KVT* __restrict__ kc,
                                                       KVT* __restrict__ vc, int Lp, int nrows, int H, int d, int Lmax) {
  __shared__ float tile[64][65];
  const int h = blockIdx.x, b = blockIdx.y, r0 = blockIdx.z * 64;
  const size_t bh = (size_t)b * H + h;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int rr = i >> 6, c = i & 63;
    const int r = r0 + rr;
    float kval = 0.f;
    if (r < nrows) {
      const size_t base = ((size_t)b * Lp + r) * 3 * d + h * 64 + c;
      kval = __bfloat162float(__ushort_as_bfloat16(hi[base + d])) + (lo ? __bfloat162float(__ushort_as_bfloat16(lo[base + d])) : 0.f);
      const float vval = __bfloat162float(__ushort_as_bfloat16(hi[base + 2 * d])) + (lo ? __bfloat162float(__ushort_as_bfloat16(lo[base + 2 * d])) : 0.f);
      kv_store(vc + (bh * Lmax + r) * 64 + c, vval);
    }
    tile[rr][c] = kval;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int c = i >> 6, rr = i & 63;
    if (r0 + rr < nrows) kv_store(kc + k_index(bh, c, r0 + rr, Lmax), tile[rr][c]);
  }
}

constexpr int DEC_MAXL = 2560;
constexpr int DEC_MAX_SPLIT = 32;     // key-range splits per (batch, head)
constexpr int DEC_CHUNK = 128;        // keys per CTA: K^T slab (64 x 128) + V slab (128 x 64) staged in shared memory
constexpr int DEC_WS = 68;            // floats per partial: m, l, pad, pad, o[64]

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// grid (heads, batch, splits), 256 threads.  Each CTA owns a contiguous range of <= DEC_CHUNK keys.  One thread issues bulk
// async copies (cp.async.bulk -> mbarrier) for the K^T slab (64 row segments) and the V slab (one contiguous block) of that range
// at kernel entry, so the whole KV traffic of the CTA is in flight while the other threads finish q (and k/v of the newest key,
// appended to the cache by the CTA that owns it) from the split-K partials of the QKV GEMM.  Scores (thread = key, camera-bias row
// added BEFORE the 1/sqrt(d_head) scale), block softmax and P.V (thread = channel) then run out of shared memory; the CTA leaves an
// (m, l, o[64]) partial and the last CTA to arrive per (batch, head) merges them: x1 = y + concat_heads(softmax(...) V)
// (flash-decoding with a fused combine).  Optionally the last head of a batch row applies LayerNorm (ln2) to the finished row.
// CK = keys per CTA: 128 (fp32 cache: 64 KB of slabs) or 256 (2-byte caches: the same 64 KB, half as many CTAs, every thread owns a key).
template <typename KVT, int CK>
__global__ void __launch_bounds__(256) dec_attn_kernel(const float* __restrict__ qkv_part, int ks, long long zstride,
                                                       const float* __restrict__ bqkv, const float* __restrict__ y,
                                                       const float* __restrict__ bias, int bias_ld, KVT* __restrict__ kc,
                                                       KVT* __restrict__ vc, float* __restrict__ x1, const int* __restrict__ step_ptr,
                                                       float* __restrict__ ws, unsigned int* __restrict__ counters, int nc, int H, int d,
                                                       int Lmax, float scale, unsigned int* __restrict__ row_counters,
                                                       const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float ln_eps,
                                                       uint16_t* __restrict__ ln_hi, uint16_t* __restrict__ ln_lo,
                                                       const uint8_t* __restrict__ layout, int lay_blk, int lay_ld) {
  extern __shared__ __align__(128) uint8_t dsm[];
  KVT* Ks = reinterpret_cast<KVT*>(dsm);                                  // [CK / 128 cache blocks][64][128]
  KVT* Vs = reinterpret_cast<KVT*>(dsm + 64 * CK * sizeof(KVT));          // [keys][64]
  __shared__ float q[64], knew[64], vnew[64], red[8], sc[CK];
  __shared__ float opart[4][64];
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned int ticket;
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, S = gridDim.z, tid = threadIdx.x;
  pdl_launch_dependents();
  // Under programmatic dependent launch the code up to pdl_wait() overlaps the previous kernel (the QKV GEMM of this layer).  It only
  // reads the step counter and the KV cache, both last written many kernels ago (every earlier kernel of the chain has completed once
  // the immediate predecessor runs), so the whole K/V slab of the CTA is already in flight while the GEMM drains.
  const int r = nc + *step_ptr - 1;
  const int n = r + 1;
  const int j0 = sp * CK;
  const int cnt = max(0, min(n - j0, CK));                 // valid keys of this CTA (key r included if in range)
  const int pitch = (cnt + 7) & ~7;                        // copied keys: 16-byte granules; stale tail entries are never used
  const bool owns_new = (r >= j0 && r < j0 + CK);
  const size_t bh = (size_t)b * H + h;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    if (cnt > 0) {
      const uint32_t kblocks = (uint32_t)((cnt + DEC_CHUNK - 1) / DEC_CHUNK);          // 128-key cache blocks touched (consecutive in memory)
      const uint32_t kbytes = kblocks * 64u * DEC_CHUNK * (uint32_t)sizeof(KVT), vbytes = (uint32_t)pitch * 64u * (uint32_t)sizeof(KVT);
      mbar_expect_tx(&bar, kbytes + vbytes);
      bulk_g2s(Ks, kc + k_index(bh, 0, j0, Lmax), kbytes, &bar);           // whole 128-key K^T blocks: one contiguous copy
      bulk_g2s(Vs, vc + (bh * Lmax + j0) * 64, vbytes, &bar);
    }
  }
  pdl_wait();
  if (tid < 192) {
    const int which = tid >> 6, c = tid & 63;
    if (cnt > 0 && (which == 0 || owns_new)) {
      const int col = which * d + h * 64 + c;
      const float* pp = qkv_part + (size_t)b * 3 * d + col;
      float v = __ldg(bqkv + col);
      int z = 0;
      for (; z + 4 <= ks; z += 4) v += (pp[z * zstride] + pp[(z + 1) * zstride]) + (pp[(z + 2) * zstride] + pp[(z + 3) * zstride]);
      for (; z < ks; ++z) v += pp[z * zstride];
      if (which == 0) q[c] = v;
      else if (which == 1) { knew[c] = v; kv_store(kc + k_index(bh, c, r, Lmax), v); }
      else { vnew[c] = v; kv_store(vc + (bh * Lmax + r) * 64 + c, v); }
    }
  }
  __syncthreads();
  float m = -INFINITY, sum = 0.f;
  if (cnt > 0) {
    mbar_wait(&bar, 0);
    if (owns_new && tid < 64) {        // the slab may hold a stale copy of the newest key: take it from registers instead
      kv_store(Ks + (((r - j0) >> 7) * 64 + tid) * DEC_CHUNK + ((r - j0) & 127), knew[tid]);
      kv_store(Vs + (size_t)(r - j0) * 64 + tid, vnew[tid]);
    }
    __syncthreads();
    const float* brow = bias ? bias + (size_t)r * bias_ld + j0 : nullptr;
    if (tid < cnt) {
      float d0 = 0.f, d1 = 0.f;
      const KVT* kcol = Ks + (tid >> 7) * 64 * DEC_CHUNK + (tid & 127);       // this key's column inside its cache block
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        d0 = fmaf(q[c], kv_load(kcol + c * DEC_CHUNK), d0);
        d1 = fmaf(q[c + 1], kv_load(kcol + (c + 1) * DEC_CHUNK), d1);
      }
      m = ((d0 + d1) + (brow ? brow[tid] : 0.f)) * scale;
      // per-head block layout (density < 1): key block (j0 + tid) / blk of query block r / blk absent -> not attended
      if (layout != nullptr && !layout[((size_t)h * lay_ld + r / lay_blk) * lay_ld + (j0 + tid) / lay_blk]) m = -INFINITY;
    }
    const float mloc = block_max_256(m, red);
    float e = 0.f;
    if (tid < cnt) { e = (m == -INFINITY) ? 0.f : expf(m - mloc); sc[tid] = e; }
    sum = block_sum_256(e, red);       // its barriers publish sc[]
    m = mloc;
    const int g = tid >> 6, c = tid & 63;
    float a0 = 0.f, a1 = 0.f;
    int jj = g;
    for (; jj + 4 < cnt; jj += 8) {
      a0 = fmaf(sc[jj], kv_load(Vs + (size_t)jj * 64 + c), a0);
      a1 = fmaf(sc[jj + 4], kv_load(Vs + (size_t)(jj + 4) * 64 + c), a1);
    }
    if (jj < cnt) a0 = fmaf(sc[jj], kv_load(Vs + (size_t)jj * 64 + c), a0);
    opart[g][c] = a0 + a1;
  } else if (tid < 256) {
    opart[tid >> 6][tid & 63] = 0.f;
  }
  __syncthreads();
  float* wp = ws + (bh * S + sp) * DEC_WS;
  if (tid < 64) {
    wp[4 + tid] = (opart[0][tid] + opart[1][tid]) + (opart[2][tid] + opart[3][tid]);
    if (tid == 0) { wp[0] = m; wp[1] = sum; }
  }
  // publish the partial: the CTA barrier orders the 64 writers before thread 0, whose single gpu-scope fence + atomic releases them
  // (fence cumulativity) - one MEMBAR per CTA instead of one per warp, which was 10 % of the kernel's stall samples
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    ticket = atomicAdd(&counters[bh], 1u);
    __threadfence();                       // acquire side for the last CTA: the other CTAs' partials are visible after the barrier below
  }
  __syncthreads();
  if (ticket != (unsigned)(S - 1)) return;
  if (tid < 64) {
    const volatile float* wv = ws + bh * S * DEC_WS;
    float M = -INFINITY;
    for (int i = 0; i < S; ++i) M = fmaxf(M, wv[i * DEC_WS]);
    float Lsum = 0.f, o = 0.f;
    for (int i = 0; i < S; ++i) {
      const float mi = wv[i * DEC_WS];
      const float w = (mi == -INFINITY) ? 0.f : expf(mi - M);
      Lsum += wv[i * DEC_WS + 1] * w;
      o += wv[i * DEC_WS + 4 + tid] * w;
    }
    const size_t idx = (size_t)b * d + h * 64 + tid;
    x1[idx] = y[idx] + o / Lsum;
    if (tid == 0) counters[bh] = 0u;      // self-reset for the next launch
  }
  if (ln_gamma == nullptr) return;
  // ---- fused LayerNorm (ln2) of the finished row: the last head of batch row b normalises x1[b,:] into the MLP operand planes
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    ticket = atomicAdd(&row_counters[b], 1u);
    __threadfence();
  }
  __syncthreads();
  if (ticket != (unsigned)(H - 1)) return;
  const bool act = tid < (d >> 2);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) v = __ldcg(reinterpret_cast<const float4*>(x1 + (size_t)b * d) + tid);
  const float mean = block_sum_256(act ? (v.x + v.y) + (v.z + v.w) : 0.f, red) / d;
  v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
  const float var = block_sum_256(act ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f, red) / d;
  const float rstd = rsqrtf(var + ln_eps);
  if (tid == 0) row_counters[b] = 0u;
  if (!act) return;
  const float4 g = __ldg(reinterpret_cast<const float4*>(ln_gamma) + tid), be = __ldg(reinterpret_cast<const float4*>(ln_beta) + tid);
  const float4 o4 = make_float4(v.x * rstd * g.x + be.x, v.y * rstd * g.y + be.y, v.z * rstd * g.z + be.z, v.w * rstd * g.w + be.w);
  __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
  split_bf16(o4.x, h0, l0); split_bf16(o4.y, h1, l1); split_bf16(o4.z, h2, l2); split_bf16(o4.w, h3, l3);
  reinterpret_cast<uint2*>(ln_hi + (size_t)b * d)[tid] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
  if (ln_lo != nullptr) reinterpret_cast<uint2*>(ln_lo + (size_t)b * d)[tid] = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
}

// ------------------------------------------------------------------------------------------------
// Sampling tail (cond_transformer_multi_view.py:138-142,200-219): logits/T, top-k filter keeping ties with the k-th
// value, softmax, multinomial (Philox4x32-10 stream keyed by (seed, step, batch)) or greedy argmax; writes the token
// into the (cam,h,w) token grid at forward_shuffle_idx[step].  One CTA of 256 threads per batch element; V <= 4096.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr int SAMPLE_MAXV = 4096;

__global__ void __launch_bounds__(256) dec_sample_kernel(const float* __restrict__ part, int ks, long long zstride, int vpad, int V,
                                                         float inv_temperature, int top_k, int greedy, unsigned long long seed,
                                                         const long long* __restrict__ forced, const int* __restrict__ fwd,
                                                         long long* __restrict__ cam_idx, long long* __restrict__ tokens_out,
                                                         float* __restrict__ trace, float* __restrict__ probs_out,
                                                         const int* __restrict__ step_ptr, int n_img, int hw, int ncam) {
  __shared__ float lg[SAMPLE_MAXV];
  __shared__ float srt[SAMPLE_MAXV];
  __shared__ float red[8];
  __shared__ float wsum[8];
  __shared__ int found;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, s = *step_ptr;
  const int B = gridDim.x;
  int np2 = 1;
  while (np2 < V) np2 <<= 1;
  for (int i = tid; i < np2; i += 256) {
    float v = -INFINITY;
    if (i < V) {
      v = 0.f;
      for (int z = 0; z < ks; ++z) v += part[z * zstride + (size_t)b * vpad + i];
      if (trace != nullptr) trace[((size_t)s * B + b) * V + i] = v;
      v *= inv_temperature;
      lg[i] = v;
    }
    srt[i] = v;
  }
  __syncthreads();
  float thr = -INFINITY;
  if (top_k > 0 && top_k < V) {
    // bitonic sort, descending
    for (int k = 2; k <= np2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += 256) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const bool desc = ((i & k) == 0);
            const float a = srt[i], c = srt[ixj];
            if (desc ? (a < c) : (a > c)) { srt[i] = c; srt[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    thr = srt[top_k - 1];
  }
  float m = -INFINITY;
  for (int i = tid; i < V; i += 256) m = fmaxf(m, lg[i]);
  m = block_max_256(m, red);
  // probabilities of 4 consecutive entries per thread (index order, needed for the inverse-CDF scan)
  const int per = (V + 255) / 256;
  float local = 0.f;
  for (int e = 0; e < per; ++e) {
    const int i = tid * per + e;
    if (i < V) {
      const float p = (lg[i] >= thr) ? expf(lg[i] - m) : 0.f;
      lg[i] = p;
      local += p;
    }
  }
  const float total = block_sum_256(local, red);
  if (probs_out != nullptr)
    for (int i = tid; i < V; i += 256) probs_out[(size_t)b * V + i] = lg[i] / total;
  int token = 0;
  // forced[b][s] >= 0: teacher forcing / partial decoding (the token is given: ground-truth cameras of
  // cond_transformer_multi_view.py:161-165,181-182); negative entries are sampled like everything else.  Uniform per CTA.
  const long long fz = (forced != nullptr) ? forced[(size_t)b * n_img + s] : -1;
  if (fz >= 0) {
    token = (int)fz;
  } else if (greedy) {
    // argmax, lowest index on ties
    float best = -1.f;
    int bi = 0;
    for (int i = tid; i < V; i += 256)
      if (lg[i] > best) { best = lg[i]; bi = i; }
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __shared__ float wb[8];
    __shared__ int wi[8];
    if ((tid & 31) == 0) { wb[tid >> 5] = best; wi[tid >> 5] = bi; }
    __syncthreads();
    best = wb[0]; bi = wi[0];
    for (int w = 1; w < 8; ++w)
      if (wb[w] > best || (wb[w] == best && wi[w] < bi)) { best = wb[w]; bi = wi[w]; }
    token = bi;
  } else {
    uint32_t rnd[4];
    philox4x32_10((uint32_t)s, (uint32_t)b, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), rnd);
    const float u = ((rnd[0] >> 8) + 0.5f) * (1.0f / 16777216.0f) * total;   // (0, total)
    // exclusive prefix of `local` over threads
    float incl = local;
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    if (tid == 0) found = V;      // sentinel
    __syncthreads();
    float base = 0.f;
    for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
    float run = base + incl - local;
    for (int e = 0; e < per; ++e) {
      const int i = tid * per + e;
      if (i < V && lg[i] > 0.f) {
        run += lg[i];
        if (run > u) { atomicMin(&found, i); break; }
      }
    }
    __syncthreads();
    token = found;
    if (token >= V) {   // numerical edge: u landed past the last bucket -> last entry with non-zero probability
      int last = 0;
      for (int i = 0; i < V; ++i) if (lg[i] > 0.f) last = i;
      token = last;
    }
  }
  if (tid == 0) {
    const int j = fwd[s];
    cam_idx[((size_t)b * ncam + j / hw) * hw + (j % hw)] = token;
    if (tokens_out != nullptr) tokens_out[(size_t)b * n_img + s] = token;
  }
}

__global__ void dec_advance_kernel(int* step) {
  pdl_launch_dependents();
  pdl_wait();
  *step += 1;
}

// ---------------------------------------------------------------- launchers
int launch_dec_reduce_ln(const float* partials, int ks, long long zstride, const float* bias, const float* residual, long long res_stride,
                         const float* gamma, const float* beta, float eps, float* x_out, float* y, uint16_t* hi, uint16_t* lo, int rows, int d,
                         cudaStream_t st) {
  if (d % 4 != 0 || d > 1024 || rows < 1) return BEVGEN_ERR_ARG;
  if (launch_k(dec_reduce_ln_kernel, dim3(rows), dim3(256), 0, st, partials, ks, zstride, bias, residual, res_stride, gamma, beta, eps, x_out, y, hi, lo,
               d) != cudaSuccess) return BEVGEN_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
int launch_dec_reduce_act(const float* partials, int ks, long long zstride, const float* bias, uint16_t* hi, uint16_t* lo, int rows, int n,
                          int gelu, cudaStream_t st) {
  if (n % 4 != 0 || rows < 1) return BEVGEN_ERR_ARG;
  if (launch_k(dec_reduce_act_kernel, dim3((rows * (n / 4) + 255) / 256), dim3(256), 0, st, partials, ks, zstride, bias, hi, lo, rows, n, gelu) != cudaSuccess)
    return BEVGEN_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
int launch_kv_store(const uint16_t* hi, const uint16_t* lo, void* kc, void* vc, int kv_bf16, int B, int Lp, int nrows, int H, int d, int Lmax,
                    cudaStream_t st) {
  if (B < 1 || B > 65535 || nrows < 1 || (Lmax & 127)) return BEVGEN_ERR_ARG;
  dim3 grid(H, B, (nrows + 63) / 64);
  if (kv_bf16 == 2) kv_store_kernel<__half><<<grid, 256, 0, st>>>(hi, lo, (__half*)kc, (__half*)vc, Lp, nrows, H, d, Lmax);
  else if (kv_bf16) kv_store_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(hi, lo, (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, Lp, nrows, H, d, Lmax);
  else kv_store_kernel<float><<<grid, 256, 0, st>>>(hi, lo, (float*)kc, (float*)vc, Lp, nrows, H, d, Lmax);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
static inline int dec_splits(int Lmax, int ck = DEC_CHUNK) { return (Lmax + ck - 1) / ck; }
int dec_attn_workspace_floats(int B, int H) { return B * H * DEC_MAX_SPLIT * DEC_WS; }

int launch_dec_attn(const float* qkv_part, int ks, long long zstride, const float* bqkv, const float* y, const float* bias, int bias_ld,
                    void* kc, void* vc, int kv_bf16, float* x1, const int* step_ptr, float* ws, unsigned int* counters, int B, int nc, int H,
                    int d, int Lmax, float scale, unsigned int* row_counters, const float* ln_gamma, const float* ln_beta, float ln_eps,
                    uint16_t* ln_hi, uint16_t* ln_lo, const uint8_t* layout, int lay_blk, int lay_ld, int /*sm_count*/, cudaStream_t st) {
  if (Lmax > DEC_MAXL || B < 1 || B > 65535 || d != H * 64 || (Lmax & 127) || dec_splits(Lmax) > DEC_MAX_SPLIT) return BEVGEN_ERR_ARG;
  if (ln_gamma != nullptr && (!row_counters || !ln_beta || !ln_hi || d > 1024)) return BEVGEN_ERR_ARG;
  if (layout != nullptr && (lay_blk < 1 || lay_ld < 1)) return BEVGEN_ERR_ARG;
  dim3 grid(H, B, dec_splits(Lmax, kv_bf16 ? 256 : DEC_CHUNK));
  if (kv_bf16) {
    static bool configured2 = false;
    if (!configured2) {
      if (cudaFuncSetAttribute(dec_attn_kernel<__half, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 256 * 2) != cudaSuccess ||
          cudaFuncSetAttribute(dec_attn_kernel<__nv_bfloat16, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 256 * 2) != cudaSuccess)
        return BEVGEN_ERR_CUDA;
      configured2 = true;
    }
  }
  if (kv_bf16 == 2) {           // fp16 cache: half the KV bytes of the fp32 cache at ~3e-4 logit error (tests/test_decode_gpu.py)
    const int smem = 2 * 64 * 256 * 2;
    if (launch_k(dec_attn_kernel<__half, 256>, grid, dim3(256), smem, st, qkv_part, ks, zstride, bqkv, y, bias, bias_ld, (__half*)kc, (__half*)vc, x1, step_ptr,
                 ws, counters, nc, H, d, Lmax, scale, row_counters, ln_gamma, ln_beta, ln_eps, ln_hi, ln_lo, layout, lay_blk, lay_ld) != cudaSuccess) return BEVGEN_ERR_CUDA;
  } else if (kv_bf16) {
    const int smem = 2 * 64 * 256 * 2;
    if (launch_k(dec_attn_kernel<__nv_bfloat16, 256>, grid, dim3(256), smem, st, qkv_part, ks, zstride, bqkv, y, bias, bias_ld, (__nv_bfloat16*)kc,
                 (__nv_bfloat16*)vc, x1, step_ptr, ws, counters, nc, H, d, Lmax, scale, row_counters, ln_gamma, ln_beta, ln_eps, ln_hi, ln_lo, layout, lay_blk, lay_ld) !=
        cudaSuccess) return BEVGEN_ERR_CUDA;
  } else {
    const int smem = 2 * 64 * DEC_CHUNK * 4;
    static bool configured = false;
    if (!configured) {
      if (cudaFuncSetAttribute(dec_attn_kernel<float, DEC_CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return BEVGEN_ERR_CUDA;
      configured = true;
    }
    if (launch_k(dec_attn_kernel<float, DEC_CHUNK>, grid, dim3(256), smem, st, qkv_part, ks, zstride, bqkv, y, bias, bias_ld, (float*)kc, (float*)vc, x1, step_ptr, ws,
                 counters, nc, H, d, Lmax, scale, row_counters, ln_gamma, ln_beta, ln_eps, ln_hi, ln_lo, layout, lay_blk, lay_ld) != cudaSuccess) return BEVGEN_ERR_CUDA;
  }
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
int launch_dec_sample(const float* part, int ks, long long zstride, int vpad, int V, float temperature, int top_k, int greedy,
                      unsigned long long seed, const long long* forced, const int* fwd, long long* cam_idx, long long* tokens_out, float* trace,
                      float* probs_out, const int* step_ptr, int B, int n_img, int hw, int ncam, cudaStream_t st) {
  if (V > SAMPLE_MAXV || V < 1 || B < 1 || temperature <= 0.f) return BEVGEN_ERR_ARG;
  if (launch_k(dec_sample_kernel, dim3(B), dim3(256), 0, st, part, ks, zstride, vpad, V, 1.0f / temperature, top_k, greedy, seed, forced, fwd, cam_idx,
               tokens_out, trace, probs_out, step_ptr, n_img, hw, ncam) != cudaSuccess) return BEVGEN_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
int launch_dec_advance(int* step, cudaStream_t st) {
  if (launch_k(dec_advance_kernel, dim3(1), dim3(1), 0, st, step) != cudaSuccess) return BEVGEN_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
