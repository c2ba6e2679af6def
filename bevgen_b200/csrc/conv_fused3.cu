// Weight-stationary 16x16-block variant of conv_fused2.cu for the large feature maps.
// Measured on B200 (tools/conv_ablation.py block, 128->128 3x3 at 256x256 x 96): bf16 single pass 3.3-3.4 ms vs 3.8-4.1 ms for conv_fused2
// (used by the engine's bf16 mode); f16f8 4.1 vs 3.8 ms and bf16x3 on par (not used there: under the 1 kW cap those modes are bound by the
// energy of the MMA stream itself - the f16f8 MMA stream alone holds the tensor pipe at 98 % and the clock at ~1.2 GHz - so halving the
// weight traffic does not buy time).
//
// conv_fused2 re-streams every weight tile once per 128-pixel tile: 295 KB of weights per CTA and tile against 92 KB of activations and
// 64 KB of residual; the ablation in profiles/r01 (tools/conv_ablation.py) shows the fused conv running at the ~6 TB/s L2->SM ceiling with
// two thirds of those bytes being weights.  Here a CTA owns a 16x16 pixel BLOCK = two horizontally adjacent 128-pixel MMA tiles that share
// one (16+2) x (16+2) halo, and every weight stage is used for BOTH tiles: weight bytes per output pixel halve, the shared halo saves
// another 10 % of activation traffic.  To make room, channel slices are 32 wide (operand slot 41 KB = halo 324 px x 32 ch x 4 B-equivalent,
// weight stage 8 KB, SWIZZLE_64B) and each tile keeps ONE 128-column accumulator, double-buffered: 2 tiles x 2 buffers x 128 = 512 TMEM
// columns.  The f16f8 product therefore scales the fp16 operands so that all three partial products share one scale S = 2^(13+e):
//     x*w*S = (fp16(x)*2^6) * fp16(w*2^(e+7))  +  e4m3((x - x16)*2^13) * e4m3(w*2^e)  +  e4m3(x) * e4m3((w - w16)*2^(13+e))
// and the epilogue multiplies by 1/S.  fp16(x)*2^6 overflows at |x| >= 1024: callers use this kernel for GroupNorm-ed inputs only (the
// engine routes un-normalised inputs to conv_fused2, whose separate correction accumulator has no range limit).
// Everything else follows conv_fused2.cu: cluster of two CTAs (cta_group::2, M = 256 per MMA, each CTA stages half of every weight
// tile), warp roles (0 TMA, 1 MMA with an elected issuing lane, 3 peer->leader arrive forwarder, 4-7 epilogue, 8-15 operand producers),
// line-coalesced cp.async producer fetch, coalesced epilogue through a swizzled transpose with GroupNorm statistics of the output.
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {
namespace fused3 {

constexpr int CB_T = 16;                                   // block edge (pixels)
constexpr int CB_HW = CB_T + 2;                            // halo edge
constexpr int CB_HPIX = CB_HW * CB_HW;                     // 324 halo pixels
constexpr int CB_CHUNK_STRIDE = CB_HPIX * 16;              // 5184 B between 8-channel (2-byte) / 16-channel (1-byte) chunks (LBO)
constexpr int CB_ROW_STRIDE = CB_HW * 16;                  // 288 B between halo rows (SBO: next 8-pixel core-matrix group)
constexpr int CB_PLANE16 = 4 * CB_CHUNK_STRIDE;            // 20736 B: 32 channels of a 2-byte plane
constexpr int CB_PLANE8 = 2 * CB_CHUNK_STRIDE;             // 10368 B: 32 channels of a 1-byte plane
constexpr int CB_BN = 128;
constexpr int CB_W_TILE = (CB_BN / 2) * 32 * 2;            // 4 KB: this CTA's 64 output channels x 32 input channels (2-byte) of a weight tile
constexpr int CB_THREADS = 512;
constexpr int CB_PBLK = 41;                                // halo pixels per producer warp (8 x 41 >= 324)

template <int NPASS>
struct Cfg3 {
  static constexpr int NOPS = (NPASS >= 2) ? 2 : 1;
  static constexpr int A_SLOT = NOPS * CB_PLANE16;                       // 41472 / 20736
  static constexpr int W_STAGE = NOPS * CB_W_TILE;                       // 8 KB / 4 KB
  static constexpr int W_STAGES = (NPASS >= 2) ? 7 : 8;
  static constexpr int A_SLOTS = 3;                                      // operand ring: the producer -> MMA -> producer hand-over latencies
                                                                         // (~1 us per hop across the CTA pair) need two slices of slack
  static constexpr int STATS_BYTES = 4 * 64 * 8 + 16 + 4 * 4096;         // per-warp fp64 GroupNorm accumulators + 4 KB transpose buffer per warp
  static constexpr int STAGE_BYTES = 8 * 3 * 1024;                       // producer cp.async ring: 8 warps x 3 slots x 1 KB
  static constexpr int SMEM = A_SLOTS * A_SLOT + W_STAGES * W_STAGE + 1024 + 512 + STATS_BYTES + STAGE_BYTES;
};

__device__ __forceinline__ uint64_t make_sdesc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// K-major SWIZZLE_64B: 8-row x 64-byte atoms, 512 B apart (SBO); layout type 4
__device__ __forceinline__ uint64_t make_sdesc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((512u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
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
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CB_THREADS, 1) conv_fused3_kernel(const __grid_constant__ ConvFusedParams p) {
  using Cfg = Cfg3<NPASS>;
  constexpr int NOPS = Cfg::NOPS, WS = Cfg::W_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                           // 1024-aligned swizzled weight tiles first
  uint8_t* sA = sW + WS * Cfg::W_STAGE;
  constexpr int AS = Cfg::A_SLOTS;
  uint64_t* bars = (uint64_t*)(sA + AS * Cfg::A_SLOT);
  uint64_t* a_full = bars;            // AS
  uint64_t* a_empty = bars + AS;      // AS
  uint64_t* w_full = bars + 2 * AS;   // WS
  uint64_t* w_empty = w_full + WS;    // WS
  uint64_t* tfull = w_empty + WS;
  uint64_t* tempty = tfull + 2;
  uint64_t* a_fwd = tempty + 2;       // AS: peer CTA only
  uint32_t* tmem_slot = (uint32_t*)(a_fwd + AS);
  double* gsm = (double*)(bars + 64);                           // [4 warps][64] group sums / sums of squares of the current image
  uint8_t* sEpi = (uint8_t*)(((uintptr_t)(gsm + 4 * 64) + 15) & ~(uintptr_t)15);            // [4 warps][4 KB] epilogue transpose buffers
  uint8_t* sStage = sEpi + 4 * 4096;                                                          // producer staging ring

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blocks_w = (p.W + CB_T - 1) / CB_T, blocks_h = (p.H + CB_T - 1) / CB_T;
  const int n_tiles_n = (p.Cout + CB_BN - 1) / CB_BN;
  const int KC = p.Cin / 32;                                    // 32-channel slices
  // work units of a cluster: (pair of consecutive 16x16 blocks) x (128-channel output tile); CTA rank r takes block 2*pair + r.
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const long long m_blocks = (long long)p.N * blocks_w * blocks_h;
  const long long pairs = (m_blocks + 1) / 2;
  const long long total = pairs * n_tiles_n;
  const int n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;
  const long long t_begin = total * cid / n_clusters, t_end = total * (cid + 1) / n_clusters;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmW[0]);
    if (NPASS >= 2) tma_prefetch_desc(&p.tmW[1]);
  }
  if (warp == 1 && lane == 0) {
    // "empty" / tfull barriers collect one tcgen05.commit from EACH of the two MMA-issuing warps
    for (int i = 0; i < AS; ++i) { mbar_init(&a_full[i], 9); mbar_init(&a_empty[i], 2); mbar_init(&a_fwd[i], 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 2); mbar_init(&tempty[i], 8); }
    for (int i = 0; i < WS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 2); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  if (threadIdx.x >= 128 && threadIdx.x < 384) gsm[threadIdx.x - 128] = 0.0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto block_valid = [&](long long t) -> bool { return (t / n_tiles_n) * 2 + rank < m_blocks; };
  auto decode = [&](long long t, int& n, int& bh, int& bw, int& nt) {
    nt = (int)(t % n_tiles_n);
    long long r = (t / n_tiles_n) * 2 + rank;            // this CTA's block of the pair
    if (r >= m_blocks) r = m_blocks - 1;                 // odd tail: duplicate work, results discarded
    bw = (int)(r % blocks_w); r /= blocks_w;
    bh = (int)(r % blocks_h);
    n = (int)(r / blocks_h);
  };

  if (warp == 0) {
    // ===================== weight TMA =====================
    if (lane == 0 && !(p.dbg & 16)) {
      uint32_t ws = 0, wphase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        const int nt = (int)(t % n_tiles_n);
        const int row0 = nt * CB_BN + (int)rank * (CB_BN / 2);
        for (int kc = 0; kc < KC; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&w_empty[ws], wphase ^ 1);
            if (leader) mbar_expect_tx(&w_full[ws], 2 * Cfg::W_STAGE);         // both CTAs' halves are credited to the leader's barrier
#pragma unroll
            for (int o = 0; o < NOPS; ++o)
              tma_load_2d_2sm(sW + ws * Cfg::W_STAGE + o * CB_W_TILE, &p.tmW[o], &w_full[ws], kc * 32, tap * p.Cout + row0);
            if (++ws == WS) { ws = 0; wphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issue (whole warp runs the control flow, one elected lane issues) =====================
    // Two issuing warps on two different warp schedulers: warp 1 drives the block's left tile, warp 2 the right tile (separate accumulators,
    // same operands).  A single issuing warp shares its scheduler with an epilogue and two producer warps and, with ~70 instructions per
    // 8-MMA stage, fell behind the tensor pipe whenever those were busy (profiles/r01 ablation: epilogue + MMA serialised).
    if (leader) {
      const int j = warp - 1;
      const bool elected = elect_one();
      // M = 256 across the CTA pair.  NPASS == 2: A/B format field 0 = F16 under kind::f16 and = E4M3 under kind::f8f6f4 (same bits)
      const uint32_t idesc = (NPASS == 2) ? (make_idesc_bf16(256, CB_BN, 0, 0) & ~((1u << 7) | (1u << 10))) : make_idesc_bf16(256, CB_BN, 0, 0);
      const uint64_t adesc0 = make_sdesc_noswz(smem_u32(sA), CB_CHUNK_STRIDE, CB_ROW_STRIDE);
      const uint64_t bdesc0 = make_sdesc_sw64(smem_u32(sW));
      constexpr uint64_t AK = (2 * CB_CHUNK_STRIDE) >> 4;          // next 16 (2-byte) channels of A
      constexpr uint64_t AP = CB_PLANE16 >> 4, AP8 = CB_PLANE8 >> 4, WT = CB_W_TILE >> 4, T1 = (8 * 16) >> 4;   // T1: second tile = +8 pixels
      const bool issue = elected && !(p.dbg & 8);
      uint32_t as = 0, aphase = 0, ws = 0, wphase = 0, acc = 0, acc_phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * (2 * CB_BN);
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&a_full[as], aphase);
          tc_fence_after();
          const uint64_t a_slot = adesc0 + (uint64_t)(as * (Cfg::A_SLOT >> 4));
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (!(p.dbg & 16)) mbar_wait(&w_full[ws], wphase);
            tc_fence_after();
            const uint64_t at = a_slot + (uint64_t)((tap / 3) * CB_HW + (tap % 3));     // shifted window inside the halo (16-byte units)
            const uint64_t bt = bdesc0 + (uint64_t)(ws * (Cfg::W_STAGE >> 4));
            if (issue) {
              const uint32_t first = (kc == 0 && tap == 0) ? 0u : 1u;
              {                                              // the two 128-pixel tiles of the block share this weight stage
                const uint32_t d = d0 + j * CB_BN;
                const uint64_t aj = at + j * T1;
                if (NPASS == 2) {
                  umma_f16_2sm(d, aj, bt, idesc, first);
                  umma_f16_2sm(d, aj + AK, bt + 2, idesc, 1u);
                  umma_f8_2sm(d, aj + AP, bt + WT, idesc, 1u);                 // e4m3((x - x16) 2^13) * e4m3(w 2^e)
                  umma_f8_2sm(d, aj + AP + AP8, bt + WT + 2, idesc, 1u);        // e4m3(x) * e4m3((w - w16) 2^(13+e))
                } else if (NPASS == 3) {
#pragma unroll
                  for (int k = 0; k < 2; ++k) {
                    umma_f16_2sm(d, aj + AP + k * AK, bt + k * 2, idesc, (k == 0) ? first : 1u);     // small terms first
                    umma_f16_2sm(d, aj + k * AK, bt + WT + k * 2, idesc, 1u);
                    umma_f16_2sm(d, aj + k * AK, bt + k * 2, idesc, 1u);
                  }
                } else {
                  umma_f16_2sm(d, aj, bt, idesc, first);
                  umma_f16_2sm(d, aj + AK, bt + 2, idesc, 1u);
                }
              }
            }
            if (elected) umma_commit_2sm(&w_empty[ws]);
            if (++ws == WS) { ws = 0; wphase ^= 1; }
          }
          if (elected) umma_commit_2sm(&a_empty[as]);
          if (++as == AS) { as = 0; aphase ^= 1; }
        }
        if (elected) umma_commit_2sm(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // peer CTA: one cluster-scope release per operand slot instead of one per producer warp
    if (lane == 0 && !leader) {
      const long long n_slots = (t_end - t_begin) * KC;
      uint32_t as = 0, aphase = 0;
      for (long long i = 0; i < n_slots; ++i) {
        mbar_wait(&a_fwd[as], aphase);
        mbar_arrive_cluster(mapa_u32(smem_u32(&a_full[as]), 0));
        if (++as == AS) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue =====================
    const int q = warp - 4;
    const int et = threadIdx.x - 128;
    const int unit = lane & 7, rsub = lane >> 3;
    const uint32_t est = smem_u32(sEpi) + q * 4096;      // this warp's 4 KB transpose buffer
    const int cpg = p.Cout / 32;                         // channels per GroupNorm group (4, 8, 16 or 32)
    const int cpg_log2 = 31 - __clz(cpg);
    int roff[8];                                          // element offset of transposed step i from the tile's first pixel (tile-invariant)
    uint32_t ld_off[8];                                   // swizzled shared-memory offset of that step's 16-byte unit
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      roff[i] = ((i >> 1) * p.W + 4 * (i & 1)) * p.Cout;
      const int rrow = 4 * i + rsub;
      ld_off[i] = rrow * 128 + ((unit ^ (rrow & 7)) * 16);
    }
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
      int n, bh, bw, nt;
      decode(t, n, bh, bw, nt);
      if (p.gn_sums != nullptr && n != cur_n) { flush(cur_n); cur_n = n; }
      const int n0 = nt * CB_BN;
      const bool bv = block_valid(t);
      if (p.residual != nullptr && t + 1 < t_end) {          // pull the NEXT block's residual rows into L2 while this block is processed
        int n2, bh2, bw2, nt2;
        decode(t + 1, n2, bh2, bw2, nt2);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int pr = q * 32 + lane, pw_ = bw2 * CB_T + j * 8 + (pr & 7), ph_ = bh2 * CB_T + (pr >> 3);
          if (pw_ < p.W && ph_ < p.H) {
            const float* pp = p.residual + (((long long)n2 * p.H + ph_) * p.W + pw_) * p.Cout + nt2 * CB_BN;
#pragma unroll
            for (int c4 = 0; c4 < CB_BN / 32; ++c4)
              if (nt2 * CB_BN + c4 * 32 < p.Cout) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + c4 * 32));
          }
        }
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {                          // the block's two 128-pixel tiles (left / right 8 columns)
        // Transposed domain: in step i (0..7) lane l handles tile row q*32 + 4*i + (l >> 3), columns 4*(l & 7) .. +3 of the 32-column chunk,
        // i.e. output pixel (oh0 + i/2, ow0 + 4*(i & 1)): every global load / store of the warp moves four complete 128-byte lines.
        const int ow0 = bw * CB_T + j * 8 + rsub, oh0 = bh * CB_T + q * 4;
        uint32_t okm = 0;
        if (bv) {
#pragma unroll
          for (int i = 0; i < 8; ++i) okm |= ((oh0 + (i >> 1) < p.H) && (ow0 + 4 * (i & 1) < p.W)) ? (1u << i) : 0u;
        }
        const long long obase = (((long long)n * p.H + oh0) * p.W + ow0) * p.Cout + n0 + unit * 4;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (2 * CB_BN) + j * CB_BN;
#pragma unroll 1
        for (int c = 0; c < CB_BN; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c, r);
          tmem_ld_wait();
          if (j == 1 && c + 32 >= CB_BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                   // accumulators drained: tell the leader's MMA thread
              if (leader) mbar_arrive(&tempty[acc]);
              else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
            }
          }
          const int col0 = n0 + c;
          if (col0 >= p.Cout || (p.dbg & 4)) continue;    // warp-uniform
          const float* rptr = p.residual + obase + c;
          float* optr = p.out + obase + c;
          float4 rres[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            rres[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.residual != nullptr) rres[i] = *reinterpret_cast<const float4*>(((okm >> i) & 1) ? rptr + roff[i] : p.residual);
          }
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + unit * 4));
#pragma unroll
          for (int u = 0; u < 8; ++u)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(est + lane * 128 + ((u ^ (lane & 7)) * 16)), "r"(r[4 * u]),
                         "r"(r[4 * u + 1]), "r"(r[4 * u + 2]), "r"(r[4 * u + 3]) : "memory");
          __syncwarp();
          float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
          const float osc = p.lo_scale;                         // 1 / S for the scaled f16f8 product, 1 otherwise
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(est + ld_off[i]));
            if ((okm >> i) & 1) {
              x.x = fmaf(x.x, osc, b4.x + rres[i].x); x.y = fmaf(x.y, osc, b4.y + rres[i].y);
              x.z = fmaf(x.z, osc, b4.z + rres[i].z); x.w = fmaf(x.w, osc, b4.w + rres[i].w);
              *reinterpret_cast<float4*>(optr + roff[i]) = x;
              s0 += x.x + x.y; q0 = fmaf(x.x, x.x, fmaf(x.y, x.y, q0));
              s1 += x.z + x.w; q1 = fmaf(x.z, x.z, fmaf(x.w, x.w, q1));
            }
          }
          __syncwarp();
          if (p.gn_sums != nullptr) {
            s0 += s1; q0 += q1;
            s0 += __shfl_xor_sync(0xffffffffu, s0, 8);  q0 += __shfl_xor_sync(0xffffffffu, q0, 8);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
            const int upg = cpg >> 2;                            // units per group: 1, 2, 4 or 8
            for (int w = 1; w < upg; w <<= 1) {
              s0 += __shfl_xor_sync(0xffffffffu, s0, w);
              q0 += __shfl_xor_sync(0xffffffffu, q0, w);
            }
            if (lane < 8 && (lane & (upg - 1)) == 0) {
              double* wacc = gsm + q * 64 + ((col0 >> cpg_log2) + (lane >> (cpg_log2 - 2))) * 2;
              wacc[0] += (double)s0;
              wacc[1] += (double)q0;
            }
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.gn_sums != nullptr) flush(cur_n);
  }
  else if (warp >= 8) {
    // ===================== operand producers =====================
    // Producer warp pw owns halo pixels pw*41 .. pw*41+40 of the block's 18x18 halo (6 rounds of 8 pixels); a pixel's 32-channel slice is
    // ONE 128-byte line.  fetch: instruction s of a round covers pixels 4s .. 4s+3 with lane -> (pixel l >> 3, 16-byte piece l & 7), i.e.
    // four complete lines; consume: lane -> (8-channel chunk l >> 3, pixel l & 7) reads its 32 bytes back from the warp's staging slot
    // (pieces XOR-swizzled by pixel).  See conv_fused2.cu for the protocol; slot = round % 3, ring runs across slice / block boundaries.
    const int pw = warp - 8;
    const int Hs = p.up2 ? p.H / 2 : p.H, Ws = p.up2 ? p.W / 2 : p.W;
    const int ush = p.up2 ? 1 : 0;
    constexpr int IPT = 6;
    const uint32_t stg = smem_u32(sStage) + pw * 3072;
    int fgeo[IPT][2], cgeo[IPT];
    uint32_t coff[IPT];
#pragma unroll
    for (int jj = 0; jj < IPT; ++jj) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int pi = jj * 8 + s2 * 4 + (lane >> 3), hp = pw * CB_PBLK + pi;
        fgeo[jj][s2] = (pi < CB_PBLK && hp < CB_HPIX) ? ((hp / CB_HW) | ((hp % CB_HW) << 8)) : -1;
      }
      const int pi = jj * 8 + (lane & 7), hp = pw * CB_PBLK + pi;
      cgeo[jj] = (pi < CB_PBLK && hp < CB_HPIX) ? ((hp / CB_HW) | ((hp % CB_HW) << 8)) : -1;
      coff[jj] = hp * 16;
    }
    const int cch = lane >> 3;                            // 8-channel chunk (0..3) of the 32-channel slice this lane transforms
    const uint32_t f_dst0 = stg + (lane >> 3) * 128 + (((lane & 7) ^ (lane >> 3)) * 16);
    const uint32_t f_dst1 = stg + (4 + (lane >> 3)) * 128 + (((lane & 7) ^ (4 + (lane >> 3))) * 16);
    const uint32_t c_src0 = stg + (lane & 7) * 128 + (((2 * cch) ^ (lane & 7)) * 16);
    const uint32_t c_src1 = stg + (lane & 7) * 128 + (((2 * cch + 1) ^ (lane & 7)) * 16);
    const uint32_t item_off = cch * CB_CHUNK_STRIDE;
    const uint32_t item_off8 = (cch >> 1) * CB_CHUNK_STRIDE + (cch & 1) * 8;
    const float* xw = p.x + (lane & 7) * 4;

    auto fetch = [&](int jj, bool live, int n_, int bh_, int bw_, int kc_) {
      if (live && !(p.dbg & 1)) {
        const float* base = xw + (size_t)n_ * Hs * Ws * p.Cin + kc_ * 32;
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int ge = fgeo[jj][s2];
          const int gh = bh_ * CB_T - 1 + (ge & 0xff), gw = bw_ * CB_T - 1 + (ge >> 8);
          if (ge >= 0 && (unsigned)gh < (unsigned)p.H && (unsigned)gw < (unsigned)p.W) {
            const float* src = base + (size_t)((gh >> ush) * Ws + (gw >> ush)) * p.Cin;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((s2 ? f_dst1 : f_dst0) + (jj % 3) * 1024), "l"(src) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int n = 0, bh = 0, bw = 0, nt = 0;
    if (t_begin < t_end) decode(t_begin, n, bh, bw, nt);
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) fetch(jj, t_begin < t_end, n, bh, bw, 0);
    uint32_t as = 0, aphase = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      int n2 = n, bh2 = bh, bw2 = bw, nt2 = nt;
      const bool more = t + 1 < t_end;
      if (more) decode(t + 1, n2, bh2, bw2, nt2);
      for (int kc = 0; kc < KC; ++kc) {
        const bool last_kc = (kc + 1 == KC);
        const bool nx_live = !last_kc || more;
        const int nx_n = last_kc ? n2 : n, nx_bh = last_kc ? bh2 : bh, nx_bw = last_kc ? bw2 : bw, nx_kc = last_kc ? 0 : kc + 1;
        float sc[8], sf[8];
        if (p.affine != nullptr) {
          const float4* ap = reinterpret_cast<const float4*>(p.affine + ((size_t)n * p.Cin + kc * 32 + cch * 8) * 2);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 q4 = __ldg(ap + e);
            sc[2 * e] = q4.x; sf[2 * e] = q4.y; sc[2 * e + 1] = q4.z; sf[2 * e + 1] = q4.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { sc[e] = 1.f; sf[e] = 0.f; }
        }
        mbar_wait(&a_empty[as], aphase ^ 1);
        const uint32_t dst = smem_u32(sA) + as * Cfg::A_SLOT;
#pragma unroll
        for (int jj = 0; jj < IPT; ++jj) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");          // round jj has landed (ring depth 3)
          __syncwarp();
          const int ge = cgeo[jj];
          float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          const int gh = bh * CB_T - 1 + (ge & 0xff), gw = bw * CB_T - 1 + (ge >> 8);
          const bool inside = ge >= 0 && (unsigned)gh < (unsigned)p.H && (unsigned)gw < (unsigned)p.W;
          if (inside && !(p.dbg & 2)) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(c_src0 + (jj % 3) * 1024));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(c_src1 + (jj % 3) * 1024));
          }
          __syncwarp();
          if (jj < 3) fetch(jj + 3, true, n, bh, bw, kc);          // refill the slot just consumed with the round 3 ahead
          else fetch(jj - 3, nx_live, nx_n, nx_bh, nx_bw, nx_kc);
          if (ge >= 0 && !(p.dbg & 2)) {
            if (inside) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[e], sf[e]);
              if (p.swish) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float ex, rc;
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(f[e] * -1.4426950408889634f));
                  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
                  f[e] *= rc;
                }
              }
            }
            const uint32_t o16 = dst + item_off + coff[jj];
            if (NPASS == 2) {
              // x16s = fp16(x * 2^6) (its 2^6 pairs with the weight plane's 2^(e+7)), lo8 = e4m3((x - x16s 2^-6) 2^13), x8 = e4m3(x)
              uint32_t h16[4];
              uint32_t l8[2] = {0u, 0u}, x8[2] = {0u, 0u};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h2 = __floats2half2_rn(f[2 * e] * 64.0f, f[2 * e + 1] * 64.0f);
                h16[e] = *reinterpret_cast<const uint32_t*>(&h2);
                const float2 hf = __half22float2(h2);
                const uint32_t lo2 = __nv_cvt_float2_to_fp8x2(make_float2(fmaf(hf.x, -128.0f, f[2 * e] * 8192.0f), fmaf(hf.y, -128.0f, f[2 * e + 1] * 8192.0f)),
                                                              __NV_SATFINITE, __NV_E4M3);
                const uint32_t xx2 = __nv_cvt_float2_to_fp8x2(make_float2(f[2 * e], f[2 * e + 1]), __NV_SATFINITE, __NV_E4M3);
                l8[e >> 1] |= lo2 << (16 * (e & 1));
                x8[e >> 1] |= xx2 << (16 * (e & 1));
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o16), "r"(h16[0]), "r"(h16[1]), "r"(h16[2]), "r"(h16[3]) : "memory");
              const uint32_t o8 = dst + CB_PLANE16 + item_off8 + coff[jj];
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(o8), "r"(l8[0]), "r"(l8[1]) : "memory");
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(o8 + CB_PLANE8), "r"(x8[0]), "r"(x8[1]) : "memory");
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
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o16), "r"(hh[0]), "r"(hh[1]), "r"(hh[2]), "r"(hh[3]) : "memory");
              if (NPASS == 3)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o16 + CB_PLANE16), "r"(ll[0]), "r"(ll[1]), "r"(ll[2]), "r"(ll[3]) : "memory");
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&a_full[as]);
          else mbar_arrive(&a_fwd[as]);
        }
        if (++as == AS) { as = 0; aphase ^= 1; }
      }
      n = n2; bh = bh2; bw = bw2; nt = nt2;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

template <int NPASS>
static int launch_t(const ConvFusedParams& p, int sm_count, cudaStream_t st) {
  using Cfg = Cfg3<NPASS>;
  auto kern = conv_fused3_kernel<NPASS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return BEVGEN_ERR_CUDA;
    configured = true;
  }
  const long long m_blocks = (long long)p.N * ((p.W + CB_T - 1) / CB_T) * ((p.H + CB_T - 1) / CB_T);
  const long long total = ((m_blocks + 1) / 2) * ((p.Cout + CB_BN - 1) / CB_BN);       // cluster work units
  const int max_clusters = sm_count / 2;
  const int grid = 2 * (int)(total < max_clusters ? total : max_clusters);
  kern<<<grid, CB_THREADS, Cfg::SMEM, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace fused3

int launch_conv_fused3(const ConvFusedParams& p, int npass, int sm_count, cudaStream_t st) {
  if (p.Cin % 32 != 0 || p.Cout % 32 != 0) return BEVGEN_ERR_ARG;
  if (p.up2 && ((p.H | p.W) & 1)) return BEVGEN_ERR_ARG;
  if (p.gn_sums != nullptr && (p.Cout / 32 < 4 || p.Cout / 32 > 32 || 32 % (p.Cout / 32) != 0 || p.Cout > 1024)) return BEVGEN_ERR_ARG;
  if (npass == 2) return fused3::launch_t<2>(p, sm_count, st);
  return npass == 3 ? fused3::launch_t<3>(p, sm_count, st) : fused3::launch_t<1>(p, sm_count, st);
}

}  // namespace bevgen
