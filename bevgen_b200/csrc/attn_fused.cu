// Fused stage-2 attention forward for sm_100a (replaces DeepSpeed's sdd -> bias add -> block-sparse softmax -> dsd chain,
// modules/transformer/sparse_self_attention.py:153-176, and the scores tensor it materialises):
//   O[b,i,h,:] = sum_j softmax_j( d_head^-1/2 * (Q_i.K_j + bias[i][j]) ) V_j  over allowed(i,j) = j < n_cond || (i >= n_cond && j <= i)
//   x1 = y + concat_heads(O)                                  (Block.forward residual, mingpt_sparse.py:250)
// One CTA per (batch, head, 128-query tile); flash-style loop over 128-key tiles, fully-masked tiles skipped.
//   warp 0      TMA producer: Q once, K/V tiles through a 2-stage ring, all boxes (64 cols x 128 rows) of the fused qkv planes
//   warp 1      tcgen05 issuer: S = Q.K^T into a double-buffered TMEM tile, PV = P.V (V read MN-major, no transpose) into a second
//   warp 2      TMEM allocator
//   warps 4-19  softmax: thread = (query row, 32-key quarter): tcgen05.ld S, + fp16 bias row, base-2 online softmax, P -> swizzled smem
//               as the A operand of the PV MMA, running O kept in registers (rescaled on the fly), final normalise + residual store.
// NPASS = 3 keeps fp32-equivalent accuracy (bf16x3 split products for both MMAs, P split into hi/lo); NPASS = 1 is plain bf16.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

constexpr int AT_BM = 128, AT_BN = 128, AT_DH = 64;
constexpr int AT_TILE = AT_BM * AT_DH * 2;       // 16 KB: one (128 x 64) bf16 tile
constexpr int AT_PTILE = AT_BM * AT_BN * 2;      // 32 KB: P (128 x 128) bf16
constexpr int AT_NPART = 4;                       // softmax threads per query row: each owns 128 / 4 = 32 keys and 64 / 4 = 16 output channels
constexpr int AT_KPT = AT_BN / AT_NPART, AT_CPT = AT_DH / AT_NPART;
constexpr int AT_THREADS = 128 + 128 * AT_NPART;  // 4 control warps + 16 softmax warps

template <int NPASS>
struct AttnCfg {
  static constexpr int NOPS = (NPASS == 3) ? 2 : 1;
  static constexpr int Q_BYTES = NOPS * AT_TILE;
  static constexpr int KV_STAGE = 2 * NOPS * AT_TILE;           // K + V
  static constexpr int P_BYTES = NOPS * AT_PTILE;
  static constexpr int ALIGN_SLACK = (NPASS == 3) ? 768 : 1024;          // NPASS 3 sits 256 B under the 227 KB limit: the kernel traps if the
                                                                         // dynamic window is less than 256-byte aligned (it starts 1 KB-aligned)
  static constexpr int SMEM = Q_BYTES + 2 * KV_STAGE + P_BYTES + ALIGN_SLACK + 256 + 128 * AT_NPART * 4;
};

struct AttnParams {
  CUtensorMap tm[2];        // hi, lo planes of qkv viewed as [B*L rows][3d cols], box (64, 128)
  const uint4* bias;        // tiled, pre-scaled camera bias (see bevgen_attn_fused_fwd) or null
  const float* y;           // [B][L][d] residual input or null
  float* x1;                // [B][L][d] fp32 output or null
  uint16_t* o_hi;           // bf16 hi / lo output planes [B][L][d] or null (operand of a following GEMM: the MaskGit to_out Linear)
  uint16_t* o_lo;
  int B, H, L, nc, d;
  float scale_log2e;        // d_head^-1/2 * log2(e)
  const unsigned long long* layout64;   // block-sparse layouts (density < 1) or null: [H][L/128 query tiles][L/128 key tiles], bit (8*rb + kb)
                                       // = the (16-row, 16-key) sub-block (rb, kb) of the tile is in the head's layout; 0 = tile skipped
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// 2^x on the MUFU unit (~2 ulp, flushes denormal results): one instruction instead of exp2f's range-handling sequence
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// (v0, v1) -> packed bf16 pair hi (one cvt) and the packed pair of the remainders lo = v - hi
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - h0, v1 - h1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int NPASS>
__global__ void __launch_bounds__(AT_THREADS, 1) attn_fused_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<NPASS>;
  constexpr int NOPS = Cfg::NOPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  if ((int)(smem - smem_raw) > Cfg::ALIGN_SLACK) __trap();
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::Q_BYTES;
  uint8_t* sP = sKV + 2 * Cfg::KV_STAGE;
  uint64_t* bars = (uint64_t*)(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* v_full = bars + 3;        // 2
  uint64_t* v_empty = bars + 5;       // 2: V stage free once P.V of its tile completed
  uint64_t* k_empty = bars + 16;      // 2: K stage free as soon as Q.K^T of its tile completed (lets the next K load start a whole softmax period early)
  uint64_t* s_full = bars + 7;        // 2
  uint64_t* s_empty = bars + 9;       // 2
  uint64_t* p_full = bars + 11;       // 1
  uint64_t* pv_done = bars + 12;      // 2
  uint64_t* pv_empty = bars + 14;     // 2
  uint32_t* tmem_slot = (uint32_t*)(bars + 18);
  float* xmax = (float*)(bars + 32);  // [128 rows][AT_NPART]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = p.L / AT_BM;
  // (batch, head)-major order so the K/V of a pair stay L2-resident while its query tiles run; heavy (late) query tiles first
  const int bh = (int)(blockIdx.x / nq);
  const int qt = nq - 1 - (int)(blockIdx.x % nq);
  const int b = bh / p.H, h = bh % p.H;
  const int m0 = qt * AT_BM;
  const int T = (m0 < p.nc) ? (p.nc / AT_BN) : (m0 / AT_BN + 1);
  const int row0 = b * p.L;           // first row of this batch element in the [B*L, 3d] view
  // key tiles this CTA visits: inside the [cond | causal] support and (with layouts) holding at least one block of the head's layout.
  // Every role walks the same bit list; pipeline stages and phases follow the visit counter `it`, addresses the tile index t.
  const unsigned long long* lay = p.layout64 ? p.layout64 + ((size_t)h * nq + qt) * (p.L / AT_BN) : nullptr;
  uint32_t tiles = T >= 32 ? 0xffffffffu : ((1u << T) - 1u);
  if (lay != nullptr) {
    uint32_t m = 0;
    for (int t = 0; t < T; ++t) m |= (lay[t] != 0 ? 1u : 0u) << t;
    tiles = m;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm[0]);
    if (NPASS == 3) tma_prefetch_desc(&p.tm[1]);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4 * AT_NPART);
      mbar_init(&pv_done[i], 1); mbar_init(&pv_empty[i], 4 * AT_NPART);
    }
    mbar_init(p_full, 4 * AT_NPART);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS[2] = {tmem, tmem + 128}, tPV[2] = {tmem + 256, tmem + 320};

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int o = 0; o < NOPS; ++o) tma_load_2d(sQ + o * AT_TILE, &p.tm[o], q_full, h * AT_DH, row0 + m0);
      int it = 0;
      for (uint32_t m = tiles; m != 0; m &= m - 1, ++it) {
        const int t = __ffs(m) - 1;
        const int s = it & 1;
        uint8_t* st = sKV + s * Cfg::KV_STAGE;
        mbar_wait(&k_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[s], NOPS * AT_TILE);
#pragma unroll
        for (int o = 0; o < NOPS; ++o) tma_load_2d(st + o * AT_TILE, &p.tm[o], &k_full[s], p.d + h * AT_DH, row0 + t * AT_BN);
        mbar_wait(&v_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[s], NOPS * AT_TILE);
#pragma unroll
        for (int o = 0; o < NOPS; ++o) tma_load_2d(st + (NOPS + o) * AT_TILE, &p.tm[o], &v_full[s], 2 * p.d + h * AT_DH, row0 + t * AT_BN);
      }
    }
  } else if (warp == 1) {
    // The whole warp runs the warp-uniform control flow (barrier addresses and descriptors stay in uniform registers) and one elected
    // lane issues the MMAs / commits; descriptors are a constant high word + (address >> 4), so per-k variants are plain increments.
    // (A lane-0-only branch built each of the 36 descriptors per key tile through R2UR chains and made the issuing thread the limiter.)
    {
      uint32_t el;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
      const bool elected = el != 0;
      const uint32_t idesc_s = make_idesc_bf16(AT_BM, AT_BN, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(AT_BM, AT_DH, 0, 1);
      const uint64_t qd = make_sdesc_sw128(smem_u32(sQ), 16, 1024);               // K-major: k-step = +32 B
      const uint64_t kd0 = make_sdesc_sw128(smem_u32(sKV), 16, 1024);
      const uint64_t pd = make_sdesc_sw128(smem_u32(sP), 16, 1024);
      const uint64_t vd0 = make_sdesc_sw128(smem_u32(sKV) + NOPS * AT_TILE, 8192, 1024);   // MN-major V: k-step = 16 rows = +2048 B
      constexpr uint64_t LO = AT_TILE >> 4, PLO = AT_PTILE >> 4, STG = Cfg::KV_STAGE >> 4;
      auto issue_s = [&](int it) {                 // stages / phases depend on the visit counter only
        const int s = it & 1;
        mbar_wait(&k_full[s], (it >> 1) & 1);
        mbar_wait(&s_empty[s], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elected) {
          const uint64_t kd = kd0 + (uint64_t)s * STG;
#pragma unroll
          for (int k = 0; k < AT_DH / 16; ++k) {
            if (NPASS == 3) {
              umma_bf16(tS[s], qd + LO + k * 2, kd + k * 2, idesc_s, k == 0 ? 0u : 1u);
              umma_bf16(tS[s], qd + k * 2, kd + LO + k * 2, idesc_s, 1u);
              umma_bf16(tS[s], qd + k * 2, kd + k * 2, idesc_s, 1u);
            } else {
              umma_bf16(tS[s], qd + k * 2, kd + k * 2, idesc_s, k == 0 ? 0u : 1u);
            }
          }
          umma_commit(&s_full[s]);
          umma_commit(&k_empty[s]);
        }
      };
      const int NT = __popc(tiles);
      mbar_wait(q_full, 0);
      if (NT > 0) issue_s(0);
      for (int it = 0; it < NT; ++it) {
        if (it + 1 < NT) issue_s(it + 1);
        const int s = it & 1;
        mbar_wait(p_full, it & 1);
        mbar_wait(&v_full[s], (it >> 1) & 1);
        mbar_wait(&pv_empty[s], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elected) {
          const uint64_t vd = vd0 + (uint64_t)s * STG;
#pragma unroll
          for (int k = 0; k < AT_BN / 16; ++k) {
            const uint64_t pa = pd + (uint64_t)((k >> 2) * (AT_TILE >> 4) + (k & 3) * 2);      // two 64-key chunks of 128 rows x 128 B
            const uint64_t va = vd + (uint64_t)(k * (2048 >> 4));
            if (NPASS == 3) {
              umma_bf16(tPV[s], pa + PLO, va, idesc_pv, k == 0 ? 0u : 1u);
              umma_bf16(tPV[s], pa, va + LO, idesc_pv, 1u);
              umma_bf16(tPV[s], pa, va, idesc_pv, 1u);
            } else {
              umma_bf16(tPV[s], pa, va, idesc_pv, k == 0 ? 0u : 1u);
            }
          }
          umma_commit(&pv_done[s]);
          umma_commit(&v_empty[s]);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / output warps =====================
    // 16 warps: thread = (query row, quarter `part` of the 128-key tile / of the 64 output channels).  (With 8 warps and 64 keys per thread
    // the per-thread instruction stream - ~1000 instructions per key tile at two warps per scheduler - was the limiter of the kernel.)
    const int sw = warp - 4;
    const int quarter = warp & 3;                  // TMEM lane quarter accessible by this warp (warp index % 4)
    const int part = sw >> 2;                      // which 32 keys of the 128-key tile / which 16 of the 64 output channels
    const int row = quarter * 32 + lane;           // query row within the tile
    const int gi = m0 + row;                       // sequence position
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float o_acc[AT_CPT];
#pragma unroll
    for (int c = 0; c < AT_CPT; ++c) o_acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, corr_saved = 0.f;
    // bias tile layout [q tile][key tile][part][16-byte unit u][row][8 x fp16], values already multiplied by scale * log2(e): lanes of a
    // warp (consecutive rows) read consecutive 16-byte pieces, i.e. every load instruction moves four complete 128-byte lines (a row-major
    // [L][L] table made each instruction touch 32 lines through the 0-KB L1 of this kernel: 12 % of all stall samples in profiles/r01b)
    const int nkt = p.L / AT_BN;
    const uint4* brow = p.bias ? p.bias + ((size_t)(m0 / AT_BM) * nkt * AT_NPART + part) * (AT_KPT / 8) * AT_BM + row : nullptr;
    float* xrow = xmax + row * AT_NPART;

    int it = 0;
    for (uint32_t tm = tiles; tm != 0; tm &= tm - 1, ++it) {
      const int t = __ffs(tm) - 1;
      const int s = it & 1;
      // bias row chunk: 32 fp16 = 64 B, issued before waiting on the MMA
      uint4 bq[AT_KPT / 8];
      if (brow != nullptr) {
        const uint4* bp = brow + (size_t)t * AT_NPART * (AT_KPT / 8) * AT_BM;
#pragma unroll
        for (int u = 0; u < AT_KPT / 8; ++u) bq[u] = __ldg(bp + u * AT_BM);
      }
      mbar_wait(&s_full[s], (it >> 1) & 1);
      tc_fence_after();
      float tv[AT_KPT];
      {
        uint32_t r0[32];
        tmem_ld_32x32(tS[s] + lane_off + part * AT_KPT, r0);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) tv[j] = __uint_as_float(r0[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);
      const bool diag = (m0 >= p.nc) && (t == T - 1);       // the only partially masked tile: n0 == m0, allowed iff key <= query
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < AT_KPT / 8; ++u) {
        const __half2* hb = reinterpret_cast<const __half2*>(&bq[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = u * 8 + e * 2;
          float2 bf = brow ? __half22float2(hb[e]) : make_float2(0.f, 0.f);
          const float a0 = fmaf(tv[j], p.scale_log2e, bf.x), a1 = fmaf(tv[j + 1], p.scale_log2e, bf.y);
          tv[j] = a0; tv[j + 1] = a1;
          mx = fmaxf(mx, fmaxf(a0, a1));
        }
      }
      if (diag) {                                            // warp-uniform: only the last tile of an image-row CTA
        mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < AT_KPT; ++j) {
          if (part * AT_KPT + j > row) tv[j] = -INFINITY;
          mx = fmaxf(mx, tv[j]);
        }
      }
      if (lay != nullptr) {        // this row's 16-row block x this thread's two 16-key blocks
        const uint32_t two = (uint32_t)(lay[t] >> (((row >> 4) << 3) + part * 2)) & 3u;
        if (two != 3u) {
          mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < AT_KPT; ++j) {
            if (!((two >> (j >> 4)) & 1u)) tv[j] = -INFINITY;
            mx = fmaxf(mx, tv[j]);
          }
        }
      }
      // exchange the row maximum with the three partner threads handling the other keys of the row
      xrow[part] = mx;
      named_bar_sync(1 + quarter, 32 * AT_NPART);
      {
        const float4 m4 = *reinterpret_cast<const float4*>(xrow);
        mx = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w));
      }
      named_bar_sync(1 + quarter, 32 * AT_NPART);           // all partners have read before any slot is rewritten
      const float m_new = fmaxf(m_run, mx);                 // finite without layouts (every row has key 0 allowed); with layouts a row's first
      const float m_sub = (m_new == -INFINITY) ? 0.f : m_new;   // visited tile may hold none of its blocks: keep exp(-inf - m) = 0 well defined
      const float corr = ex2_approx(m_run - m_sub);         // 0 on the first tile
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < AT_KPT; ++j) { tv[j] = ex2_approx(tv[j] - m_sub); psum += tv[j]; }
      l_run = l_run * corr + psum;
      m_run = m_new;
      // fold in the previous tile's P.V (also guarantees the PV MMA finished reading P from smem)
      if (it > 0) {
        const int sp = (it - 1) & 1;
        mbar_wait(&pv_done[sp], ((it - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t r[16];
        tmem_ld_32x16(tPV[sp] + lane_off + part * AT_CPT, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pv_empty[sp]);
#pragma unroll
        for (int c = 0; c < AT_CPT; ++c) o_acc[c] = o_acc[c] * corr_saved + __uint_as_float(r[c]);
      }
      corr_saved = corr;
      // P -> smem, K-major SWIZZLE_128B: row r at r*128 B inside each 64-key chunk, 16-byte unit u stored at u ^ (r & 7);
      // this thread's 32 keys are units 4*(part & 1) .. +3 of chunk part >> 1
      {
        uint8_t* prow = sP + (part >> 1) * AT_TILE + row * 128;
#pragma unroll
        for (int u = 0; u < AT_KPT / 8; ++u) {
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_bf16x2(tv[u * 8 + 2 * e], tv[u * 8 + 2 * e + 1], hh[e], ll[e]);
          const int us = (((part & 1) * 4 + u) ^ (row & 7)) * 16;
          *reinterpret_cast<uint4*>(prow + us) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          if (NPASS == 3) *reinterpret_cast<uint4*>(prow + AT_PTILE + us) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        }
      }
      fence_proxy_async();          // make the generic-proxy smem writes visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // last tile's P.V
    if (it > 0) {
      const int sp = (it - 1) & 1;
      mbar_wait(&pv_done[sp], ((it - 1) >> 1) & 1);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x16(tPV[sp] + lane_off + part * AT_CPT, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < AT_CPT; ++c) o_acc[c] = o_acc[c] * corr_saved + __uint_as_float(r[c]);
    }
    // total row sum over the four partners
    xrow[part] = l_run;
    named_bar_sync(1 + quarter, 32 * AT_NPART);
    const float4 l4 = *reinterpret_cast<const float4*>(xrow);
    const float lsum = (l4.x + l4.y) + (l4.z + l4.w);
    const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;          // a row without any attended key (degenerate layout) contributes nothing
    const size_t off = ((size_t)(b * p.L + gi)) * p.d + h * AT_DH + part * AT_CPT;
#pragma unroll
    for (int c = 0; c < AT_CPT; ++c) o_acc[c] *= inv;
    if (p.y != nullptr) {
      const float4* yp = reinterpret_cast<const float4*>(p.y + off);
#pragma unroll
      for (int c = 0; c < AT_CPT / 4; ++c) {
        const float4 yv = __ldg(yp + c);
        o_acc[4 * c] += yv.x; o_acc[4 * c + 1] += yv.y; o_acc[4 * c + 2] += yv.z; o_acc[4 * c + 3] += yv.w;
      }
    }
    if (p.x1 != nullptr) {
      float4* op = reinterpret_cast<float4*>(p.x1 + off);
#pragma unroll
      for (int c = 0; c < AT_CPT / 4; ++c) op[c] = make_float4(o_acc[4 * c], o_acc[4 * c + 1], o_acc[4 * c + 2], o_acc[4 * c + 3]);
    }
    if (p.o_hi != nullptr) {
      uint32_t hh[AT_CPT / 2], ll[AT_CPT / 2];
#pragma unroll
      for (int c = 0; c < AT_CPT / 2; ++c) split_bf16x2(o_acc[2 * c], o_acc[2 * c + 1], hh[c], ll[c]);
#pragma unroll
      for (int c = 0; c < AT_CPT / 8; ++c) {
        reinterpret_cast<uint4*>(p.o_hi + off)[c] = make_uint4(hh[4 * c], hh[4 * c + 1], hh[4 * c + 2], hh[4 * c + 3]);
        if (p.o_lo != nullptr) reinterpret_cast<uint4*>(p.o_lo + off)[c] = make_uint4(ll[4 * c], ll[4 * c + 1], ll[4 * c + 2], ll[4 * c + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int NPASS>
static int launch_attn(const AttnParams& p, cudaStream_t st) {
  using Cfg = AttnCfg<NPASS>;
  auto kern = attn_fused_kernel<NPASS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return BEVGEN_ERR_CUDA;
    configured = true;
  }
  kern<<<p.B * p.H * (p.L / AT_BM), AT_THREADS, Cfg::SMEM, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_attn_fused(const CUtensorMap* tm_hi, const CUtensorMap* tm_lo, const void* bias_f16, const float* y, float* x1, int B, int H, int L,
                      int nc, int d, float scale, int npass, const unsigned long long* layout64, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st) {
  if (L % AT_BM != 0 || nc % AT_BN != 0 || nc < AT_BN || nc > L || d != H * AT_DH || L / AT_BN > 32) return BEVGEN_ERR_ARG;
  AttnParams p;
  p.tm[0] = *tm_hi;
  p.tm[1] = tm_lo ? *tm_lo : *tm_hi;
  p.bias = (const uint4*)bias_f16;
  p.y = y; p.x1 = x1;
  p.B = B; p.H = H; p.L = L; p.nc = nc; p.d = d;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.layout64 = layout64;
  p.o_hi = out_hi; p.o_lo = out_lo;
  return npass == 3 ? launch_attn<3>(p, st) : launch_attn<1>(p, st);
}

}  // namespace bevgen
