// KV-cache autoregressive decode kernels (the reference has no KV cache: Net2NetTransformer.sample re-runs the full
// 1792-token forward per token, modules/stage2/cond_transformer_multi_view.py:172-219; SURVEY.md §3.4 derives the
// equivalent cached formulation from the closed-form mask).  Every kernel reads the current step from device memory so
// one CUDA graph can be replayed for all 1536 tokens.  Step s samples decode-order token s; it processes sequence row
// r = n_cond + s - 1 (the row holding token s-1, or the last cond row for s = 0) against keys 0..r.
// The weight GEMMs of a step are swap-AB tcgen05 launches (gemm_tc.cu, GF_OUT_T) that leave split-K partials
// [ks][batch][features]; the kernels here fold "sum partials + bias (+ activation / residual / LayerNorm)" into one pass.
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
template <typename T> __device__ __forceinline__ void kv_store(T* p, float v);
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

template <typename T> struct KV2;
template <> struct KV2<float> {
  static __device__ __forceinline__ float2 load(const float* p) { return *reinterpret_cast<const float2*>(p); }
};
template <> struct KV2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 load(const __nv_bfloat16* p) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
};

constexpr int DEC_STAGES = 3;
constexpr int DEC_KSPLIT = 2;         // key-range halves per (batch, head): 2 * B * H work items per launch

// Persistent decode attention: grid = #SMs, 9 warps.  A work item is one half of the key range of a (batch, head) pair; every CTA walks
// items c, c+G, ... .  Warp 8 is the producer: it streams the 128-key blocks of the CTA's items through a 3-deep shared-memory ring with
// bulk async copies (cp.async.bulk -> mbarrier; K^T block = one contiguous 32 KB piece of the blocked cache, V rows contiguous), running
// ahead across item boundaries, so ~128 KB of KV traffic per SM stays in flight.  Warps 0-7 are consumers with NO block-wide barrier in the
// block loop: warp w owns keys [16w, 16w+16) of every block and keeps its own online-softmax state (m, l, 2 channels per lane):
//   scores: lane = (key, half of the 64 channels) out of shared memory, camera-bias row added BEFORE the 1/sqrt(d_head) scale;
//   P.V: lane = channel pair, probabilities broadcast by shuffle.
// Per item: q (and k/v of the newest key, appended to the cache by the item that owns it) are finished from the split-K partials of the
// QKV GEMM; the 8 warp states are merged in shared memory, the item's (m, l, o[64]) goes to a workspace and the last item per
// (batch, head) merges the halves:  x1 = y + concat_heads(softmax(...) V).  Optionally the last head of a batch row applies LayerNorm (ln2).
template <typename KVT>
__global__ void __launch_bounds__(288, 1) dec_attn_kernel(const float* __restrict__ qkv_part, int ks, long long zstride,
                                                          const float* __restrict__ bqkv, const float* __restrict__ y,
                                                          const float* __restrict__ bias, int bias_ld, KVT* __restrict__ kc,
                                                          KVT* __restrict__ vc, float* __restrict__ x1, const int* __restrict__ step_ptr,
                                                          float* __restrict__ ws, unsigned int* __restrict__ counters, int nc, int B, int H,
                                                          int d, int Lmax, float scale, unsigned int* __restrict__ row_counters,
                                                          const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float ln_eps,
                                                          uint16_t* __restrict__ ln_hi, uint16_t* __restrict__ ln_lo) {
  extern __shared__ __align__(128) uint8_t dsm[];
  constexpr uint32_t KBYTES = 64u * DEC_CHUNK * (uint32_t)sizeof(KVT);
  __shared__ float q[64], knew[64], vnew[64], red[8];
  __shared__ float wm[8], wl[8], wo[8][64];
  __shared__ __align__(8) uint64_t full[DEC_STAGES], empty[DEC_STAGES];
  __shared__ unsigned int ticket;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = nc + *step_ptr - 1;
  const int n = r + 1;
  const int nblk = (n + DEC_CHUNK - 1) / DEC_CHUNK;          // key blocks in use
  const int bps = (nblk + DEC_KSPLIT - 1) / DEC_KSPLIT;      // blocks per item (the last item of a pair may get fewer, possibly 0)
  const int items = B * H * DEC_KSPLIT;
  const int G = gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < DEC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 8) {
    // ===================== producer =====================
    if (lane == 0) {
      int slot = 0;
      for (int item = blockIdx.x; item < items; item += G) {
        const size_t bh = item / DEC_KSPLIT;
        const int b0 = (item % DEC_KSPLIT) * bps, b1 = min(nblk, b0 + bps);
        for (int blk = b0; blk < b1; ++blk, ++slot) {
          const int st = slot % DEC_STAGES;
          mbar_wait(&empty[st], ((slot / DEC_STAGES) & 1) ^ 1);
          const int j0 = blk * DEC_CHUNK;
          const int cnt = min(n - j0, DEC_CHUNK);
          const uint32_t vbytes = (uint32_t)((cnt + 7) & ~7) * 64u * (uint32_t)sizeof(KVT);
          uint8_t* base = dsm + (size_t)st * 2 * KBYTES;
          mbar_expect_tx(&full[st], KBYTES + vbytes);
          bulk_g2s(base, kc + k_index(bh, 0, j0, Lmax), KBYTES, &full[st]);
          bulk_g2s(base + KBYTES, vc + (bh * Lmax + j0) * 64, vbytes, &full[st]);
        }
      }
    }
    return;
  }
  // ===================== consumers (warps 0-7, named barrier 1 with 256 threads) =====================
  const int kk = lane & 15, hh = lane >> 4;                  // scores: key within the warp's 16, channel half
  int slot = 0;
  for (int item = blockIdx.x; item < items; item += G) {
    const int bh = item / DEC_KSPLIT, sp = item % DEC_KSPLIT;
    const int b = bh / H, h = bh % H;
    const int b0 = sp * bps, b1 = min(nblk, b0 + bps);
    const bool owns_new = (b0 < b1) && (b1 == nblk);
    asm volatile("bar.sync 1, 256;" ::: "memory");           // previous item's q / wm / wo are no longer read
    if (b0 < b1 && tid < 192) {
      const int which = tid >> 6, c = tid & 63;
      if (which == 0 || owns_new) {
        const int col = which * d + h * 64 + c;
        const float* pp = qkv_part + (size_t)b * 3 * d + col;
        float v = __ldg(bqkv + col);
        int z = 0;
        for (; z + 4 <= ks; z += 4) v += (pp[z * zstride] + pp[(z + 1) * zstride]) + (pp[(z + 2) * zstride] + pp[(z + 3) * zstride]);
        for (; z < ks; ++z) v += pp[z * zstride];
        if (which == 0) q[c] = v;
        else if (which == 1) { knew[c] = v; kv_store(kc + k_index(bh, c, r, Lmax), v); }
        else { vnew[c] = v; kv_store(vc + ((size_t)bh * Lmax + r) * 64 + c, v); }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float qh[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) qh[c] = q[hh * 32 + c];
    float m = -INFINITY, l = 0.f;
    float2 acc = make_float2(0.f, 0.f);
    const float* brow = bias ? bias + (size_t)r * bias_ld : nullptr;
    for (int blk = b0; blk < b1; ++blk, ++slot) {
      const int st = slot % DEC_STAGES;
      const KVT* Ks = reinterpret_cast<const KVT*>(dsm + (size_t)st * 2 * KBYTES);
      const KVT* Vs = reinterpret_cast<const KVT*>(dsm + (size_t)st * 2 * KBYTES + KBYTES);
      const int j0 = blk * DEC_CHUNK;
      const int jk = warp * 16 + kk;                          // key within the block
      const int j = j0 + jk;
      mbar_wait(&full[st], (slot / DEC_STAGES) & 1);
      float sv = -INFINITY;
      if (j < n) {
        float dot = 0.f;
        if (j < r) {
#pragma unroll
          for (int c = 0; c < 32; ++c) dot = fmaf(qh[c], kv_load(Ks + (hh * 32 + c) * DEC_CHUNK + jk), dot);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) dot = fmaf(qh[c], knew[hh * 32 + c], dot);
        }
        sv = dot;
      }
      sv += __shfl_xor_sync(0xffffffffu, sv, 16);             // both channel halves; -inf stays -inf
      if (j < n) sv = (sv + (brow ? brow[j] : 0.f)) * scale;
      float tmax = sv;
#pragma unroll
      for (int o = 8; o; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      const float m_new = fmaxf(m, tmax);
      if (m_new != -INFINITY) {                               // warp-uniform: this warp has at least one valid key so far
        const float corr = (m == -INFINITY) ? 0.f : expf(m - m_new);
        const float p = (j < n) ? expf(sv - m_new) : 0.f;     // lanes 16-31 mirror lanes 0-15
        float psum = (hh == 0) ? p : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
        l = l * corr + psum;
        acc.x *= corr; acc.y *= corr;
        m = m_new;
        const int nv = min(16, n - (j0 + warp * 16));         // valid keys of this warp in the block (>= 1 here)
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float pj = __shfl_sync(0xffffffffu, p, t);
          if (t < nv) {
            const int jj = j0 + warp * 16 + t;
            float2 vv;
            if (jj < r) vv = KV2<KVT>::load(Vs + (size_t)(warp * 16 + t) * 64 + 2 * lane);
            else vv = make_float2(vnew[2 * lane], vnew[2 * lane + 1]);
            acc.x = fmaf(pj, vv.x, acc.x);
            acc.y = fmaf(pj, vv.y, acc.y);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (lane == 0) { wm[warp] = m; wl[warp] = l; }
    wo[warp][2 * lane] = acc.x;
    wo[warp][2 * lane + 1] = acc.y;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float* wp = ws + ((size_t)bh * DEC_MAX_SPLIT + sp) * DEC_WS;
    if (tid < 64) {
      float M = wm[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) M = fmaxf(M, wm[w]);
      float Ls = 0.f, o = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float f = (wm[w] == -INFINITY) ? 0.f : expf(wm[w] - M);
        Ls += wl[w] * f;
        o += wo[w][tid] * f;
      }
      wp[4 + tid] = o;
      if (tid == 0) { wp[0] = M; wp[1] = Ls; }
    }
    __threadfence();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) ticket = atomicAdd(&counters[bh], 1u);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (ticket != (unsigned)(DEC_KSPLIT - 1)) continue;
    __threadfence();
    if (tid < 64) {
      const volatile float* wv = ws + (size_t)bh * DEC_MAX_SPLIT * DEC_WS;
      float M = -INFINITY;
#pragma unroll
      for (int i = 0; i < DEC_KSPLIT; ++i) M = fmaxf(M, wv[i * DEC_WS]);
      float Lsum = 0.f, o = 0.f;
#pragma unroll
      for (int i = 0; i < DEC_KSPLIT; ++i) {
        const float mi = wv[i * DEC_WS];
        const float w = (mi == -INFINITY) ? 0.f : expf(mi - M);
        Lsum += wv[i * DEC_WS + 1] * w;
        o += wv[i * DEC_WS + 4 + tid] * w;
      }
      const size_t idx = (size_t)b * d + h * 64 + tid;
      x1[idx] = y[idx] + o / Lsum;
      if (tid == 0) counters[bh] = 0u;      // self-reset for the next launch
    }
    if (ln_gamma == nullptr) continue;
    // ---- fused LayerNorm (ln2): the last head of batch row b normalises x1[b,:] into the MLP operand planes
    __threadfence();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) ticket = atomicAdd(&row_counters[b], 1u);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (ticket != (unsigned)(H - 1)) continue;
    __threadfence();
    const bool act = tid < (d >> 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) v = __ldcg(reinterpret_cast<const float4*>(x1 + (size_t)b * d) + tid);
    float s1 = act ? (v.x + v.y) + (v.z + v.w) : 0.f;
    for (int o = 16; o; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    if (lane == 0) red[warp] = s1;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float mean = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) / d;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
    float s2 = act ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f;
    for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 0) red[warp] = s2;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float rstd = rsqrtf((((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) / d + ln_eps);
    if (tid == 0) row_counters[b] = 0u;
    if (act) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(ln_gamma) + tid), be = __ldg(reinterpret_cast<const float4*>(ln_beta) + tid);
      const float4 o4 = make_float4(v.x * rstd * g.x + be.x, v.y * rstd * g.y + be.y, v.z * rstd * g.z + be.z, v.w * rstd * g.w + be.w);
      __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
      split_bf16(o4.x, h0, l0); split_bf16(o4.y, h1, l1); split_bf16(o4.z, h2, l2); split_bf16(o4.w, h3, l3);
      reinterpret_cast<uint2*>(ln_hi + (size_t)b * d)[tid] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
      if (ln_lo != nullptr) reinterpret_cast<uint2*>(ln_lo + (size_t)b * d)[tid] = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
    }
  }
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
  if (forced != nullptr) {
    token = (int)forced[(size_t)b * n_img + s];
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

__global__ void dec_advance_kernel(int* step) { *step += 1; }

// ---------------------------------------------------------------- launchers
int launch_dec_reduce_ln(const float* partials, int ks, long long zstride, const float* bias, const float* residual, long long res_stride,
                         const float* gamma, const float* beta, float eps, float* x_out, float* y, uint16_t* hi, uint16_t* lo, int rows, int d,
                         cudaStream_t st) {
  if (d % 4 != 0 || d > 1024 || rows < 1) return BEVGEN_ERR_ARG;
  dec_reduce_ln_kernel<<<rows, 256, 0, st>>>(partials, ks, zstride, bias, residual, res_stride, gamma, beta, eps, x_out, y, hi, lo, d);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
int launch_dec_reduce_act(const float* partials, int ks, long long zstride, const float* bias, uint16_t* hi, uint16_t* lo, int rows, int n,
                          int gelu, cudaStream_t st) {
  if (n % 4 != 0 || rows < 1) return BEVGEN_ERR_ARG;
  dec_reduce_act_kernel<<<(rows * (n / 4) + 255) / 256, 256, 0, st>>>(partials, ks, zstride, bias, hi, lo, rows, n, gelu);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
int launch_kv_store(const uint16_t* hi, const uint16_t* lo, void* kc, void* vc, int kv_bf16, int B, int Lp, int nrows, int H, int d, int Lmax,
                    cudaStream_t st) {
  if (B < 1 || B > 65535 || nrows < 1 || (Lmax & 127)) return BEVGEN_ERR_ARG;
  dim3 grid(H, B, (nrows + 63) / 64);
  if (kv_bf16) kv_store_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(hi, lo, (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, Lp, nrows, H, d, Lmax);
  else kv_store_kernel<float><<<grid, 256, 0, st>>>(hi, lo, (float*)kc, (float*)vc, Lp, nrows, H, d, Lmax);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
static inline int dec_splits(int Lmax) { return (Lmax + DEC_CHUNK - 1) / DEC_CHUNK; }
int dec_attn_workspace_floats(int B, int H) { return B * H * DEC_MAX_SPLIT * DEC_WS; }

template <typename KVT>
static int launch_dec_attn_t(const float* qkv_part, int ks, long long zstride, const float* bqkv, const float* y, const float* bias, int bias_ld,
                             KVT* kc, KVT* vc, float* x1, const int* step_ptr, float* ws, unsigned int* counters, int B, int nc, int H, int d,
                             int Lmax, float scale, unsigned int* row_counters, const float* ln_gamma, const float* ln_beta, float ln_eps,
                             uint16_t* ln_hi, uint16_t* ln_lo, int sm_count, cudaStream_t st) {
  const int smem = DEC_STAGES * 2 * 64 * DEC_CHUNK * (int)sizeof(KVT);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(dec_attn_kernel<KVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return BEVGEN_ERR_CUDA;
    configured = true;
  }
  const int per_sm = (sizeof(KVT) == 2) ? 2 : 1;      // bf16 ring is 96 KB: two CTAs per SM
  dec_attn_kernel<KVT><<<sm_count * per_sm, 288, smem, st>>>(qkv_part, ks, zstride, bqkv, y, bias, bias_ld, kc, vc, x1, step_ptr, ws, counters, nc, B,
                                                             H, d, Lmax, scale, row_counters, ln_gamma, ln_beta, ln_eps, ln_hi, ln_lo);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_dec_attn(const float* qkv_part, int ks, long long zstride, const float* bqkv, const float* y, const float* bias, int bias_ld,
                    void* kc, void* vc, int kv_bf16, float* x1, const int* step_ptr, float* ws, unsigned int* counters, int B, int nc, int H,
                    int d, int Lmax, float scale, unsigned int* row_counters, const float* ln_gamma, const float* ln_beta, float ln_eps,
                    uint16_t* ln_hi, uint16_t* ln_lo, int sm_count, cudaStream_t st) {
  if (Lmax > DEC_MAXL || B < 1 || B > 65535 || d != H * 64 || (Lmax & 127) || dec_splits(Lmax) > DEC_MAX_SPLIT) return BEVGEN_ERR_ARG;
  if (ln_gamma != nullptr && (!row_counters || !ln_beta || !ln_hi || d > 1024)) return BEVGEN_ERR_ARG;
  if (kv_bf16)
    return launch_dec_attn_t<__nv_bfloat16>(qkv_part, ks, zstride, bqkv, y, bias, bias_ld, (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, x1, step_ptr, ws,
                                            counters, B, nc, H, d, Lmax, scale, row_counters, ln_gamma, ln_beta, ln_eps, ln_hi, ln_lo, sm_count, st);
  return launch_dec_attn_t<float>(qkv_part, ks, zstride, bqkv, y, bias, bias_ld, (float*)kc, (float*)vc, x1, step_ptr, ws, counters, B, nc, H, d,
                                  Lmax, scale, row_counters, ln_gamma, ln_beta, ln_eps, ln_hi, ln_lo, sm_count, st);
}
int launch_dec_sample(const float* part, int ks, long long zstride, int vpad, int V, float temperature, int top_k, int greedy,
                      unsigned long long seed, const long long* forced, const int* fwd, long long* cam_idx, long long* tokens_out, float* trace,
                      float* probs_out, const int* step_ptr, int B, int n_img, int hw, int ncam, cudaStream_t st) {
  if (V > SAMPLE_MAXV || V < 1 || B < 1 || temperature <= 0.f) return BEVGEN_ERR_ARG;
  dec_sample_kernel<<<B, 256, 0, st>>>(part, ks, zstride, vpad, V, 1.0f / temperature, top_k, greedy, seed, forced, fwd, cam_idx, tokens_out,
                                       trace, probs_out, step_ptr, n_img, hw, ncam);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}
int launch_dec_advance(int* step, cudaStream_t st) {
  dec_advance_kernel<<<1, 1, 0, st>>>(step);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
