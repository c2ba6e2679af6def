// 2-CTA (cta_group::2) variant of conv_fused.cu: a cluster of two CTAs computes two adjacent 128-pixel tiles with ONE M=256 tcgen05.mma
// per k-step issued by the leader CTA.  Each CTA stages only HALF of every weight tile (64 of the 128 output channels) in its shared
// memory, which halves both the L2->SM weight traffic (38 GB per 128->128 conv at 256x256 in profiles/r01, the limiter of the 1-CTA
// kernel at the ~6.9 TB/s L2 cap) and the shared-memory bytes read per MMA (A 4 KB + B 2 KB instead of 4 + 4 per 64 cycles).
// Protocol: both CTAs' TMA threads load their weight half and complete_tx on the LEADER's w_full barrier; both CTAs' producer warps
// arrive on the leader's a_full; the leader's tcgen05.commit multicasts to both CTAs' w_empty / a_empty / tfull; the peer's epilogue
// warps arrive remotely on the leader's tempty.
// 3x3 stride-1 "same" convolution reading the fp32 NHWC activation DIRECTLY, with the GroupNorm-apply + swish + bf16 split
// (+ optional nearest 2x upsampling) fused into the operand path: 8 producer warps load the (16+2) x (8+2) x 64 halo with coalesced
// 32-byte global loads, apply y = swish(x * scale[n,c] + shift[n,c]), split y = hi + lo and store both bf16 planes into the
// shared-memory halo layout of conv_halo.cu; the separate `prep` pass (4 B read + 4 B written per element, 24 % of a VQGAN step in
// profiles/r01) and its operand planes in HBM disappear.  Everything else is conv_halo.cu:
// gemm_tc.cu loads one shifted activation box per tap (9 boxes per 64-channel chunk), which makes shared-memory bandwidth the
// limiter of the bf16x3 product (profiles/r01: tensor pipe 65 %: per k-step the MMAs read 96 KB while TMA writes 64 KB, at 128 B/clk).
// Here the activation halo (16+2) x (8+2) pixels x 64 channels is loaded ONCE per channel chunk into a no-swizzle, 16-byte-chunk-major
// layout [8-channel chunk][halo pixel][16 B]; every tap is then just a different start address of the UMMA A descriptor
// (K-major, SWIZZLE_NONE: core matrices = 8 pixels x 8 channels, LBO = chunk stride, SBO = halo row stride), so activation smem
// write traffic drops 9x (6.4x counting the halo) and the tensor pipe becomes the limiter.  Weights stream through a 4-stage SWIZZLE_128B
// ring exactly as in gemm_tc.cu.  The epilogue fuses bias, residual and the GroupNorm(32) statistics of the OUTPUT (sum / sum of squares
// per (image, group), accumulated per CTA in shared memory and flushed with fp64 atomics when the image changes).
// Output tile = 16 rows x 8 columns of pixels x 128 output channels; CTAs own contiguous tile ranges (halo reuse in L2, few flushes).
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {
namespace fused2 {

constexpr int CH_TW = 8, CH_TH = 16;                       // output tile (pixels)
constexpr int CH_HW = CH_TW + 2, CH_HH = CH_TH + 2;        // halo
constexpr int CH_HPIX = CH_HW * CH_HH;                     // 180 halo pixels
constexpr int CH_CHUNK_STRIDE = CH_HPIX * 16;              // bytes between 8-channel chunks (LBO)
constexpr int CH_ROW_STRIDE = CH_HW * 16;                  // bytes between halo rows (SBO: next 8-pixel core-matrix group)
constexpr int CH_A_PLANE = 8 * CH_CHUNK_STRIDE;            // 23040 B: 64 channels
constexpr int CH_BN = 128;
constexpr int CH_W_TILE = (CH_BN / 2) * 64 * 2;            // 8 KB: this CTA's half (64 output channels) of a weight tile
constexpr int CH_THREADS = 512;        // warps 0-3 control, 4-7 epilogue, 8-15 operand producers

template <int NPASS>
struct ConvFusedCfg {
  static constexpr int NOPS = (NPASS >= 2) ? 2 : 1;                       // NPASS == 2: fp16 plane + packed fp8 correction planes
  static constexpr int TMEM_COLS = (NPASS == 2) ? 512 : 256;             // NPASS == 2 keeps the fp8 correction sum in its own accumulator
  static constexpr int A_SLOT = NOPS * CH_A_PLANE;                       // 46080 / 23040
  static constexpr int A_SLOT_PAD = (A_SLOT + 1023) & ~1023;
  static constexpr int W_STAGE = NOPS * CH_W_TILE;                       // 32 KB / 16 KB
  static constexpr int W_STAGES = (NPASS >= 2) ? 6 : 8;
  static constexpr int STATS_BYTES = 4 * 64 * 8 + 4 * 16 * 33 * 4 + 16;       // per-warp fp64 accumulators + transpose scratch
  static constexpr int STAGE_BYTES = 3 * 2 * 256 * 16;                   // producer cp.async ring: 3 slots x 2 halves x 256 threads x 16 B
  static constexpr int SMEM = 2 * A_SLOT_PAD + W_STAGES * W_STAGE + 1024 + 512 + STATS_BYTES + STAGE_BYTES;
};

__device__ __forceinline__ uint64_t make_sdesc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// GroupNorm partial statistics of one 32-channel chunk (CPG channels per group, CPG >= 4): every lane (= pixel) forms its 2*(32/CPG)
// per-group (sum, sum of squares), the warp transposes them through a private shared-memory scratch and lanes 0..NV-1 each reduce one
// value over the 32 pixels and add it (fp64, no atomics: each slot has a single owner lane) into the warp's accumulator row.
template <int CPG>
__device__ __forceinline__ void gn_accumulate(const float (&v)[32], bool row_ok, int lane, int col0, float* scratch /*[16][33]*/,
                                              double* wacc /*[64] of this warp*/) {
  constexpr int NG = 32 / CPG, NV = 2 * NG;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) { const float x = row_ok ? v[g * CPG + j] : 0.f; s += x; ss += x * x; }
    scratch[(2 * g) * 33 + lane] = s;
    scratch[(2 * g + 1) * 33 + lane] = ss;
  }
  __syncwarp();
  if (lane < NV) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += scratch[lane * 33 + i];
    wacc[(col0 / CPG) * 2 + lane] += (double)t;        // group (col0/CPG + lane/2), component lane&1
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit of the barrier address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same issue shape for 8-bit operands (e4m3 x e4m3 -> fp32, K = 32 per instruction)
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CH_THREADS, 1) conv_fused2_kernel(const __grid_constant__ ConvFusedParams p) {
  using Cfg = ConvFusedCfg<NPASS>;
  constexpr int NOPS = Cfg::NOPS, WS = Cfg::W_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                           // 1024-aligned swizzled weight tiles first
  uint8_t* sA = sW + WS * Cfg::W_STAGE;
  uint64_t* bars = (uint64_t*)(sA + 2 * Cfg::A_SLOT_PAD);
  uint64_t* a_full = bars;            // 2
  uint64_t* a_empty = bars + 2;       // 2
  uint64_t* w_full = bars + 4;        // WS
  uint64_t* w_empty = bars + 4 + WS;  // WS
  uint64_t* tfull = bars + 4 + 2 * WS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  double* gsm = (double*)(bars + 64);                           // [4 warps][64] group sums / sums of squares of the current image
  float* gscr = (float*)(gsm + 4 * 64);                         // [4 warps][16][33] transpose scratch
  uint8_t* sStage = (uint8_t*)(((uintptr_t)(gscr + 4 * 16 * 33) + 15) & ~(uintptr_t)15);   // producer staging ring

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (p.W + CH_TW - 1) / CH_TW, tiles_h = (p.H + CH_TH - 1) / CH_TH;
  const int n_tiles_n = (p.Cout + CH_BN - 1) / CH_BN;
  const int KC = p.Cin / 64;
  // work units of a cluster: (pair of adjacent 128-pixel tiles) x (128-channel output tile); CTA rank r takes pixel tile 2*pair + r.
  // When the number of pixel tiles is odd the last pair's second CTA recomputes the last tile and discards it (tile_valid = false).
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const long long m_tiles = (long long)p.N * tiles_w * tiles_h;
  const long long pairs = (m_tiles + 1) / 2;
  const long long total = pairs * n_tiles_n;
  const int n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;
  const long long t_begin = total * cid / n_clusters, t_end = total * (cid + 1) / n_clusters;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmW[0]);
    if (NPASS >= 2) tma_prefetch_desc(&p.tmW[1]);
  }
  if (warp == 1 && lane == 0) {
    // leader-side barriers collect arrivals from BOTH CTAs (a_full: 16 producer warps, tempty: 8 epilogue warps)
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 16); mbar_init(&a_empty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    for (int i = 0; i < WS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  if (threadIdx.x >= 128 && threadIdx.x < 384) gsm[threadIdx.x - 128] = 0.0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // barrier inits and TMEM allocation of both CTAs are visible before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto tile_valid = [&](long long t) -> bool { return (t / n_tiles_n) * 2 + rank < m_tiles; };

  auto decode = [&](long long t, int& n, int& th, int& tw, int& nt) {
    nt = (int)(t % n_tiles_n);
    long long r = (t / n_tiles_n) * 2 + rank;            // this CTA's pixel tile of the pair
    if (r >= m_tiles) r = m_tiles - 1;                   // odd tail: duplicate work, results discarded
    tw = (int)(r % tiles_w); r /= tiles_w;
    th = (int)(r % tiles_h);
    n = (int)(r / tiles_h);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t wi = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        int n, th, tw, nt;
        decode(t, n, th, tw, nt);
        for (int kc = 0; kc < KC; ++kc) {
          for (int tap = 0; tap < 9; ++tap, ++wi) {
            const int ws = wi % WS;
            mbar_wait(&w_empty[ws], ((wi / WS) & 1) ^ 1);
            if (leader) mbar_expect_tx(&w_full[ws], 2 * Cfg::W_STAGE);         // both CTAs' halves are credited to the leader's barrier
#pragma unroll
            for (int o = 0; o < NOPS; ++o)
              tma_load_2d_2sm(sW + ws * Cfg::W_STAGE + o * CH_W_TILE, &p.tmW[o], &w_full[ws], kc * 64,
                              tap * p.Cout + nt * CH_BN + (int)rank * (CH_BN / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // M = 256 across the CTA pair.  NPASS == 2: A/B format field 0 = F16 under kind::f16 and = E4M3 under kind::f8f6f4 (same bits)
      const uint32_t idesc = (NPASS == 2) ? (make_idesc_bf16(256, CH_BN, 0, 0) & ~((1u << 7) | (1u << 10))) : make_idesc_bf16(256, CH_BN, 0, 0);
      constexpr uint32_t ACC_COLS = (NPASS == 2) ? 2 * CH_BN : CH_BN;
      uint32_t ai = 0, wi = 0, acc = 0, acc_phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kc = 0; kc < KC; ++kc, ++ai) {
          const int as = ai & 1;
          mbar_wait(&a_full[as], (ai >> 1) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + as * Cfg::A_SLOT_PAD);
          for (int tap = 0; tap < 9; ++tap, ++wi) {
            const int ws = wi % WS;
            mbar_wait(&w_full[ws], (wi / WS) & 1);
            tc_fence_after();
            const uint32_t b_base = smem_u32(sW + ws * Cfg::W_STAGE);
            const uint32_t a_tap = a_base + ((tap / 3) * CH_HW + (tap % 3)) * 16;     // shifted window inside the halo
            if (p.dbg & 8) {
            } else if (NPASS == 2) {
              // x*w ~= x16*w16 (fp16 MMA, accumulator 0) + [xlo8*w8 + x8*wlo8] / S (e4m3 MMAs at twice the rate, accumulator 1)
              const uint32_t first = (kc == 0 && tap == 0) ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(d_tmem, make_sdesc_noswz(a_tap + k * 2 * CH_CHUNK_STRIDE, CH_CHUNK_STRIDE, CH_ROW_STRIDE),
                              make_sdesc_sw128(b_base + k * 32, 16, 1024), idesc, (k == 0) ? first : 1u);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                umma_f8_2sm(d_tmem + CH_BN, make_sdesc_noswz(a_tap + CH_A_PLANE + k * 2 * CH_CHUNK_STRIDE, CH_CHUNK_STRIDE, CH_ROW_STRIDE),
                            make_sdesc_sw128(b_base + CH_W_TILE + k * 32, 16, 1024), idesc, (k == 0) ? first : 1u);
                umma_f8_2sm(d_tmem + CH_BN, make_sdesc_noswz(a_tap + CH_A_PLANE + CH_A_PLANE / 2 + k * 2 * CH_CHUNK_STRIDE, CH_CHUNK_STRIDE, CH_ROW_STRIDE),
                            make_sdesc_sw128(b_base + CH_W_TILE + 64 + k * 32, 16, 1024), idesc, 1u);
              }
            } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_hi = make_sdesc_noswz(a_tap + k * 2 * CH_CHUNK_STRIDE, CH_CHUNK_STRIDE, CH_ROW_STRIDE);
              const uint64_t b_hi = make_sdesc_sw128(b_base + k * 32, 16, 1024);
              const uint32_t first = (kc == 0 && tap == 0 && k == 0) ? 0u : 1u;
              if (NPASS == 3) {
                const uint64_t a_lo = make_sdesc_noswz(a_tap + CH_A_PLANE + k * 2 * CH_CHUNK_STRIDE, CH_CHUNK_STRIDE, CH_ROW_STRIDE);
                const uint64_t b_lo = make_sdesc_sw128(b_base + CH_W_TILE + k * 32, 16, 1024);
                umma_bf16_2sm(d_tmem, a_lo, b_hi, idesc, first);
                umma_bf16_2sm(d_tmem, a_hi, b_lo, idesc, 1u);
                umma_bf16_2sm(d_tmem, a_hi, b_hi, idesc, 1u);
              } else {
                umma_bf16_2sm(d_tmem, a_hi, b_hi, idesc, first);
              }
            }
            }
            umma_commit_2sm(&w_empty[ws]);
          }
          umma_commit_2sm(&a_empty[as]);
        }
        umma_commit_2sm(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const int cpg = p.Cout / 32;                        // channels per GroupNorm group (4, 8 or 16 here)
    uint32_t acc = 0, acc_phase = 0;
    int cur_n = -1;
    auto flush = [&](int n_img) {                         // all 128 epilogue threads
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (n_img >= 0 && et < 64) {
        const double v = (gsm[et] + gsm[64 + et]) + (gsm[128 + et] + gsm[192 + et]);
        if (v != 0.0) atomicAdd(p.gn_sums + (size_t)n_img * 64 + et, v);
        gsm[et] = 0.0; gsm[64 + et] = 0.0; gsm[128 + et] = 0.0; gsm[192 + et] = 0.0;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    for (long long t = t_begin; t < t_end; ++t) {
      int n, th, tw, nt;
      decode(t, n, th, tw, nt);
      if (p.gn_sums != nullptr && n != cur_n) { flush(cur_n); cur_n = n; }
      const int n0 = nt * CH_BN;
      const int ow = tw * CH_TW + (row % CH_TW), oh = th * CH_TH + (row / CH_TW);
      const bool row_ok = (ow < p.W) && (oh < p.H) && tile_valid(t);
      const long long roff = (((long long)n * p.H + oh) * p.W + ow) * p.Cout + n0;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ((NPASS == 2) ? 2 * CH_BN : CH_BN);
#pragma unroll 1
      for (int c = 0; c < CH_BN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        if (NPASS == 2) {                                    // add the scaled fp8 correction accumulator
          uint32_t r2[32];
          tmem_ld_32x32(taddr + CH_BN + c, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(r2[j]), p.lo_scale, __uint_as_float(r[j])));
        }
        tmem_ld_wait();
        if (c + 32 >= CH_BN) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {                                   // accumulator drained: tell the leader's MMA thread
            if (leader) mbar_arrive(&tempty[acc]);
            else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
          }
        }
        const int col0 = n0 + c;
        if (col0 >= p.Cout || (p.dbg & 4)) continue;    // warp-uniform
        float v[32];
        {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);      // col0 % 32 == 0: 16-byte aligned, warp-uniform
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(bp + j);
            v[4 * j] = __uint_as_float(r[4 * j]) + b4.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
          }
        }
        if (row_ok) {
          if (p.residual != nullptr) {
            const float4* rp = reinterpret_cast<const float4*>(p.residual + roff + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 rv = rp[j];
              v[4 * j] += rv.x; v[4 * j + 1] += rv.y; v[4 * j + 2] += rv.z; v[4 * j + 3] += rv.w;
            }
          }
          float4* op = reinterpret_cast<float4*>(p.out + roff + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (p.gn_sums != nullptr) {
          switch (cpg) {
            case 4: gn_accumulate<4>(v, row_ok, lane, col0, gscr + q * 16 * 33, gsm + q * 64); break;
            case 8: gn_accumulate<8>(v, row_ok, lane, col0, gscr + q * 16 * 33, gsm + q * 64); break;
            case 16: gn_accumulate<16>(v, row_ok, lane, col0, gscr + q * 16 * 33, gsm + q * 64); break;
            default: gn_accumulate<32>(v, row_ok, lane, col0, gscr + q * 16 * 33, gsm + q * 64); break;
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.gn_sums != nullptr) flush(cur_n);
  }
  else if (warp >= 8) {
    // ===================== operand producers: global fp32 -> affine (+swish) -> bf16 hi/lo -> halo in shared memory =====================
    // Every thread owns items (8-channel chunk, halo pixel) j*256 + pt of each (tile, channel chunk): 1440 items = 6 per thread, pixel
    // fastest (conflict-free 16-byte shared stores).  The raw 32 bytes of an item are fetched with cp.async into a PRIVATE 3-deep ring
    // of staging slots, 3 items ahead of the one being transformed (across tile / chunk boundaries), so global latency is hidden without
    // holding the data in registers.
    const int pt = threadIdx.x - 256;                    // 0..255
    const int Hs = p.up2 ? p.H / 2 : p.H, Ws = p.up2 ? p.W / 2 : p.W;      // source geometry
    constexpr int DEPTH = 3, IPT = 6;                     // ring depth, items per thread per (tile, kc)
    uint8_t* stg = sStage + pt * 16;                      // slot s, half h at stg + (s*2 + h) * 256*16

    // fetch cursor (runs DEPTH items ahead of the transform cursor); advanced incrementally, no divisions by runtime values
    long long f_t = t_begin;
    int f_kc = 0, f_j = 0, f_slot = 0, f_n = 0, f_th = 0, f_tw = 0, f_nt = 0;
    if (t_begin < t_end) decode(f_t, f_n, f_th, f_tw, f_nt);
    auto fetch = [&]() {
      if (f_t < t_end) {
        const int i = pt + 256 * f_j;
        if (i < 8 * CH_HPIX) {
          const int chunk = i / CH_HPIX, px = i % CH_HPIX;
          const int gh = f_th * CH_TH - 1 + px / CH_HW, gw = f_tw * CH_TW - 1 + px % CH_HW;
          if (gh >= 0 && gh < p.H && gw >= 0 && gw < p.W && !(p.dbg & 1)) {
            const int sh = p.up2 ? gh >> 1 : gh, sw_ = p.up2 ? gw >> 1 : gw;
            const float* src = p.x + (((size_t)f_n * Hs + sh) * Ws + sw_) * p.Cin + f_kc * 64 + chunk * 8;
            const uint32_t d0 = smem_u32(stg + (f_slot * 2) * 256 * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 256 * 16), "l"(src + 4) : "memory");
          }
        }
        if (++f_j == IPT) {
          f_j = 0;
          if (++f_kc == KC) {
            f_kc = 0;
            if (++f_t < t_end) decode(f_t, f_n, f_th, f_tw, f_nt);
          }
        }
      }
      if (++f_slot == DEPTH) f_slot = 0;
      asm volatile("cp.async.commit_group;" ::: "memory");      // one group per item (possibly empty) keeps the wait arithmetic uniform
    };

    for (int g = 0; g < DEPTH; ++g) fetch();
    int c_slot = 0;                                       // transform cursor's staging slot
    uint32_t ai = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      int n, th, tw, nt;
      decode(t, n, th, tw, nt);
      for (int kc = 0; kc < KC; ++kc, ++ai) {
        const int as = ai & 1;
        mbar_wait(&a_empty[as], ((ai >> 1) & 1) ^ 1);
        uint8_t* dst = sA + as * Cfg::A_SLOT_PAD;
#pragma unroll 1
        for (int jj = 0; jj < IPT; ++jj) {
          asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");     // item g has landed
          const int i = pt + 256 * jj;
          if (i < 8 * CH_HPIX && !(p.dbg & 2)) {
            const int chunk = i / CH_HPIX, px = i % CH_HPIX;
            const int gh = th * CH_TH - 1 + px / CH_HW, gw = tw * CH_TW - 1 + px % CH_HW;
            const int off = chunk * CH_CHUNK_STRIDE + px * 16;
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (gh >= 0 && gh < p.H && gw >= 0 && gw < p.W) {        // zero padding applies AFTER the transform
              const uint8_t* sp = stg + (c_slot * 2) * 256 * 16;
              const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 256 * 16);
              f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
              if (p.affine != nullptr) {
                const float4* ap = reinterpret_cast<const float4*>(p.affine + ((size_t)n * p.Cin + kc * 64 + chunk * 8) * 2);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float4 sc = __ldg(ap + e);             // (scale, shift) x 2 channels
                  f[2 * e] = fmaf(f[2 * e], sc.x, sc.y);
                  f[2 * e + 1] = fmaf(f[2 * e + 1], sc.z, sc.w);
                }
              }
              if (p.swish) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __fdividef(f[e], 1.0f + __expf(-f[e]));     // swish, fast intrinsics (~1e-6 rel.)
              }
            }
            if (NPASS == 2) {
              // fp16 plane + two e4m3 planes: lo8 = (x - fp16(x)) * 2^13 and x8 = x * 4 (satfinite: out-of-range values degrade gracefully)
              uint32_t h16[4];
              uint32_t l8[2] = {0u, 0u}, x8[2] = {0u, 0u};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h2 = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                h16[e] = *reinterpret_cast<const uint32_t*>(&h2);
                const float2 hf = __half22float2(h2);
                const uint32_t lo2 = __nv_cvt_float2_to_fp8x2(make_float2((f[2 * e] - hf.x) * 8192.0f, (f[2 * e + 1] - hf.y) * 8192.0f), __NV_SATFINITE, __NV_E4M3);
                const uint32_t xx2 = __nv_cvt_float2_to_fp8x2(make_float2(f[2 * e] * 4.0f, f[2 * e + 1] * 4.0f), __NV_SATFINITE, __NV_E4M3);
                l8[e >> 1] |= lo2 << (16 * (e & 1));
                x8[e >> 1] |= xx2 << (16 * (e & 1));
              }
              *reinterpret_cast<uint4*>(dst + off) = make_uint4(h16[0], h16[1], h16[2], h16[3]);
              const int off8 = (chunk >> 1) * CH_CHUNK_STRIDE + px * 16 + (chunk & 1) * 8;      // 16-channel chunks of 1-byte elements
              *reinterpret_cast<uint2*>(dst + CH_A_PLANE + off8) = make_uint2(l8[0], l8[1]);
              *reinterpret_cast<uint2*>(dst + CH_A_PLANE + CH_A_PLANE / 2 + off8) = make_uint2(x8[0], x8[1]);
            } else {
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(f[2 * e], h0, l0);
              split_bf16(f[2 * e + 1], h1, l1);
              hh[e] = pack_bf16(h0, h1);
              ll[e] = pack_bf16(l0, l1);
            }
            *reinterpret_cast<uint4*>(dst + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            if (NPASS == 3) *reinterpret_cast<uint4*>(dst + CH_A_PLANE + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
          }
          if (++c_slot == DEPTH) c_slot = 0;
          fetch();                          // refill the slot just consumed
        }
        fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {                                     // operand slot filled: tell the leader's MMA thread
          if (leader) mbar_arrive(&a_full[as]);
          else mbar_arrive_cluster(mapa_u32(smem_u32(&a_full[as]), 0));
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer may still multicast into / arrive on this CTA's shared memory until both are done
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
}

template <int NPASS>
static int launch_conv_fused2_t(const ConvFusedParams& p, int sm_count, cudaStream_t st) {
  using Cfg = ConvFusedCfg<NPASS>;
  auto kern = conv_fused2_kernel<NPASS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return BEVGEN_ERR_CUDA;
    configured = true;
  }
  const long long m_tiles = (long long)p.N * ((p.W + CH_TW - 1) / CH_TW) * ((p.H + CH_TH - 1) / CH_TH);
  const long long total = ((m_tiles + 1) / 2) * ((p.Cout + CH_BN - 1) / CH_BN);       // cluster work units
  const int max_clusters = sm_count / 2;
  const int grid = 2 * (int)(total < max_clusters ? total : max_clusters);
  kern<<<grid, CH_THREADS, Cfg::SMEM, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_conv_fused(const ConvFusedParams& p, int npass, int sm_count, cudaStream_t st) {
  if (p.Cin % 64 != 0 || p.Cout % 32 != 0 || p.Cout % 4 != 0) return BEVGEN_ERR_ARG;
  if (p.up2 && ((p.H | p.W) & 1)) return BEVGEN_ERR_ARG;
  if (p.gn_sums != nullptr && (p.Cout / 32 < 4 || p.Cout / 32 > 32 || 32 % (p.Cout / 32) != 0 || p.Cout > 1024)) return BEVGEN_ERR_ARG;
  if (npass == 2) return launch_conv_fused2_t<2>(p, sm_count, st);
  return npass == 3 ? launch_conv_fused2_t<3>(p, sm_count, st) : launch_conv_fused2_t<1>(p, sm_count, st);
}

}  // namespace fused2

int launch_conv_fused2(const ConvFusedParams& p, int npass, int sm_count, cudaStream_t st) { return fused2::launch_conv_fused(p, npass, sm_count, st); }

}  // namespace bevgen
