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
  static constexpr int W_STAGES = (NPASS >= 2) ? 5 : 8;
  static constexpr int STATS_BYTES = 4 * 64 * 8 + 16 + 4 * 4096;             // per-warp fp64 GroupNorm accumulators + 4 KB epilogue transpose buffer per warp
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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
  uint64_t* a_fwd = tempty + 2;       // 2: peer CTA only, its 8 producer warps; warp 3 forwards each completion to the leader's a_full
  uint32_t* tmem_slot = (uint32_t*)(a_fwd + 2);
  double* gsm = (double*)(bars + 64);                           // [4 warps][64] group sums / sums of squares of the current image
  uint8_t* sEpi = (uint8_t*)(((uintptr_t)(gsm + 4 * 64) + 15) & ~(uintptr_t)15);            // [4 warps][4 KB] epilogue transpose buffers
  uint8_t* sStage = sEpi + 4 * 4096;                                                          // producer staging ring

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
    // leader-side barriers collect arrivals from BOTH CTAs (a_full: 8 own producer warps + the peer's forwarder, tempty: 8 epilogue warps)
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 9); mbar_init(&a_empty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); mbar_init(&a_fwd[i], 8);
    }
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
    if (lane == 0 && !(p.dbg & 16)) {
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
    if (leader) {
      // The WHOLE warp runs the (warp-uniform) control flow, so ptxas keeps barrier addresses / descriptors in uniform registers; one
      // elected lane issues the MMAs and commits (a lane-0-only branch made every descriptor a chain of R2UR moves and the single issuing
      // thread the limiter of the kernel: ~980 cycles per 8-MMA stage against 512 cycles of tensor time, profiles/r01 ablation).
      const bool elected = elect_one();
      // M = 256 across the CTA pair.  NPASS == 2: A/B format field 0 = F16 under kind::f16 and = E4M3 under kind::f8f6f4 (same bits)
      const uint32_t idesc = (NPASS == 2) ? (make_idesc_bf16(256, CH_BN, 0, 0) & ~((1u << 7) | (1u << 10))) : make_idesc_bf16(256, CH_BN, 0, 0);
      constexpr uint32_t ACC_COLS = (NPASS == 2) ? 2 * CH_BN : CH_BN;
      // descriptors = constant high word + (address >> 4) in the low 14 bits: all per-tap / per-k variants are small constant increments
      const uint64_t adesc0 = make_sdesc_noswz(smem_u32(sA), CH_CHUNK_STRIDE, CH_ROW_STRIDE);
      const uint64_t bdesc0 = make_sdesc_sw128(smem_u32(sW), 16, 1024);
      constexpr uint64_t AK = (2 * CH_CHUNK_STRIDE) >> 4, AP = CH_A_PLANE >> 4, WT = CH_W_TILE >> 4;
      const bool issue = elected && !(p.dbg & 8);
      uint32_t ai = 0, ws = 0, wphase = 0, acc = 0, acc_phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kc = 0; kc < KC; ++kc, ++ai) {
          const uint32_t as = ai & 1;
          mbar_wait(&a_full[as], (ai >> 1) & 1);
          tc_fence_after();
          const uint64_t a_slot = adesc0 + (uint64_t)(as * (Cfg::A_SLOT_PAD >> 4));
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (!(p.dbg & 16)) mbar_wait(&w_full[ws], wphase);
            tc_fence_after();
            const uint64_t at = a_slot + (uint64_t)((tap / 3) * CH_HW + (tap % 3));     // shifted window inside the halo (16-byte units)
            const uint64_t bt = bdesc0 + (uint64_t)(ws * (Cfg::W_STAGE >> 4));
            if (issue) {
              if (NPASS == 2) {
                // x*w ~= x16*w16 (fp16 MMA, accumulator 0) + [xlo8*w8 + x8*wlo8] / S (e4m3 MMAs at twice the rate, accumulator 1)
                const uint32_t first = (kc == 0 && tap == 0) ? 0u : 1u;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_2sm(d_tmem, at + k * AK, bt + k * 2, idesc, (k == 0) ? first : 1u);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  umma_f8_2sm(d_tmem + CH_BN, at + AP + k * AK, bt + WT + k * 2, idesc, (k == 0) ? first : 1u);
                  umma_f8_2sm(d_tmem + CH_BN, at + AP + AP / 2 + k * AK, bt + WT + 4 + k * 2, idesc, 1u);
                }
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t first = (kc == 0 && tap == 0 && k == 0) ? 0u : 1u;
                  if (NPASS == 3) {
                    umma_bf16_2sm(d_tmem, at + AP + k * AK, bt + k * 2, idesc, first);
                    umma_bf16_2sm(d_tmem, at + k * AK, bt + WT + k * 2, idesc, 1u);
                    umma_bf16_2sm(d_tmem, at + k * AK, bt + k * 2, idesc, 1u);
                  } else {
                    umma_bf16_2sm(d_tmem, at + k * AK, bt + k * 2, idesc, first);
                  }
                }
              }
            }
            if (elected) umma_commit_2sm(&w_empty[ws]);
            if (++ws == WS) { ws = 0; wphase ^= 1; }
          }
          if (elected) umma_commit_2sm(&a_empty[as]);
        }
        if (elected) umma_commit_2sm(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // peer CTA: one cluster-scope release per operand slot instead of one per producer warp (the release is a heavyweight MEMBAR.ALL.GPU)
    if (lane == 0 && !leader) {
      const long long n_slots = (t_end - t_begin) * KC;
      for (long long i = 0; i < n_slots; ++i) {
        const int as = (int)(i & 1);
        mbar_wait(&a_fwd[as], (uint32_t)(i >> 1) & 1);
        mbar_arrive_cluster(mapa_u32(smem_u32(&a_full[as]), 0));
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp - 4;
    const int et = threadIdx.x - 128;
    const int unit = lane & 7, rsub = lane >> 3;
    const uint32_t est = smem_u32(sEpi) + q * 4096;      // this warp's 4 KB transpose buffer
    const int cpg_log2 = 31 - __clz(p.Cout / 32);
    int roff[8];                                          // element offset of transposed step i from the tile's first pixel (tile-invariant)
    uint32_t ld_off[8];                                   // swizzled shared-memory offset of that step's 16-byte unit
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      roff[i] = ((i >> 1) * p.W + 4 * (i & 1)) * p.Cout;
      const int rrow = 4 * i + rsub;
      ld_off[i] = rrow * 128 + ((unit ^ (rrow & 7)) * 16);
    }
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
      // Transposed domain: in step i (0..7) lane l handles tile row q*32 + 4*i + (l >> 3), columns 4*(l & 7) .. +3 of the 32-column chunk,
      // i.e. output pixel (oh0 + i/2, ow0 + 4*(i & 1)): every global load / store of the warp moves four complete 128-byte lines.
      const int ow0 = tw * CH_TW + rsub, oh0 = th * CH_TH + q * 4;
      uint32_t okm = 0;                                      // bit i: step i lies inside the image (and the tile is not the odd-tail duplicate)
      if (tile_valid(t)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) okm |= ((oh0 + (i >> 1) < p.H) && (ow0 + 4 * (i & 1) < p.W)) ? (1u << i) : 0u;
      }
      const long long obase = (((long long)n * p.H + oh0) * p.W + ow0) * p.Cout + n0 + unit * 4;
      if (p.residual != nullptr && t + 1 < t_end) {          // pull the NEXT tile's residual rows into L2 while this tile is processed
        int n2, th2, tw2, nt2;
        decode(t + 1, n2, th2, tw2, nt2);
        const int pr = q * 32 + lane, pw_ = tw2 * CH_TW + (pr % CH_TW), ph_ = th2 * CH_TH + (pr / CH_TW);
        if (pw_ < p.W && ph_ < p.H) {
          const float* pp = p.residual + (((long long)n2 * p.H + ph_) * p.W + pw_) * p.Cout + nt2 * CH_BN;
#pragma unroll
          for (int j = 0; j < CH_BN / 32; ++j)
            if (nt2 * CH_BN + j * 32 < p.Cout) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + j * 32));
        }
      }
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
        // residual rows of the transposed domain: issued before the transpose so their latency overlaps it
        const float* rptr = p.residual + obase + c;
        float* optr = p.out + obase + c;
        float4 rres[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rres[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.residual != nullptr) rres[i] = *reinterpret_cast<const float4*>(((okm >> i) & 1) ? rptr + roff[i] : p.residual);
        }
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + unit * 4));
        // thread = row  ->  shared (16-byte units XOR-swizzled by row: conflict-free both ways)  ->  thread = (row quad, 4-column unit)
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(est + lane * 128 + ((u ^ (lane & 7)) * 16)), "r"(r[4 * u]),
                       "r"(r[4 * u + 1]), "r"(r[4 * u + 2]), "r"(r[4 * u + 3]) : "memory");
        __syncwarp();
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;          // (sum, sumsq) of channels {0,1} and {2,3} of this lane's unit
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(est + ld_off[i]));
          if ((okm >> i) & 1) {
            x.x += b4.x + rres[i].x; x.y += b4.y + rres[i].y; x.z += b4.z + rres[i].z; x.w += b4.w + rres[i].w;
            *reinterpret_cast<float4*>(optr + roff[i]) = x;
            s0 += x.x + x.y; q0 = fmaf(x.x, x.x, fmaf(x.y, x.y, q0));
            s1 += x.z + x.w; q1 = fmaf(x.z, x.z, fmaf(x.w, x.w, q1));
          }
        }
        __syncwarp();
        if (p.gn_sums != nullptr) {
          // cpg >= 4: a 4-column unit lies inside one group.  Reduce over the 4 row-quads (lanes xor 8, 16), then over the units of a group.
          s0 += s1; q0 += q1;
          s0 += __shfl_xor_sync(0xffffffffu, s0, 8);  q0 += __shfl_xor_sync(0xffffffffu, q0, 8);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
          const int upg = cpg >> 2;                            // units per group: 1, 2, 4 or 8
          for (int w = 1; w < upg; w <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, w);
            q0 += __shfl_xor_sync(0xffffffffu, q0, w);
          }
          if (lane < 8 && (lane & (upg - 1)) == 0) {
            double* wacc = gsm + q * 64 + ((col0 >> cpg_log2) + (lane >> (cpg_log2 - 2))) * 2;      // single owner lane per slot: no atomics
            wacc[0] += (double)s0;
            wacc[1] += (double)q0;
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.gn_sums != nullptr) flush(cur_n);
  }
  else if (warp >= 8) {
    // ===================== operand producers: global fp32 -> affine (+swish) -> split -> halo layout in shared memory =====================
    // Producer warp pw owns channel group g = pw & 1 (32 channels = ONE 128-byte line per pixel) of every 64-channel slice and the halo
    // pixel block pb = pw >> 1 (45 of the 180 halo pixels), processed in 6 rounds of 8 pixels.
    //   fetch:   two cp.async per lane and round; instruction s covers pixels 4s .. 4s+3 of the round with lane -> (pixel l >> 3, 16-byte
    //            piece l & 7): every instruction moves four COMPLETE 128-byte lines (the previous lane = pixel mapping asked L2 for 32
    //            half-used sectors per instruction and capped the kernel's DRAM stream at ~3 TB/s, profiles/r01 ablation);
    //   consume: lane -> (8-channel chunk l >> 3, pixel l & 7) reads its 32 bytes back from the warp's staging slot (pieces XOR-swizzled by
    //            pixel: conflict-free), so consecutive lanes write consecutive pixels of one chunk plane (conflict-free 16-byte stores).
    // Lanes exchange data through the slot, hence the __syncwarp()s around the staging reads.  All geometry is tile-invariant and lives
    // in registers; the GroupNorm affine of a lane's 8 channels is fetched once per (tile, slice); the staging ring is 3 rounds deep and
    // runs across slice / tile boundaries (slot = round % 3).
    const int pw = warp - 8;
    const int g = pw & 1, pb = pw >> 1;
    constexpr int PBLK = CH_HPIX / 4;                     // 45 halo pixels per block
    const int Hs = p.up2 ? p.H / 2 : p.H, Ws = p.up2 ? p.W / 2 : p.W;      // source geometry
    const int ush = p.up2 ? 1 : 0;
    constexpr int IPT = 6;                                // rounds per (tile, slice); ring depth 3 = IPT / 2
    const uint32_t stg = smem_u32(sStage) + pw * 3072;    // this warp's 3 slots of 1 KB: [8 pixels][8 pieces ^ pixel][16 B]
    // fetch geometry: (halo row | halo column << 8) of the pixel this lane fetches in round jj, instruction s; -1 = beyond the block
    int fgeo[IPT][2], cgeo[IPT];
    uint32_t coff[IPT];                                   // consume side: halo pixel index * 16
#pragma unroll
    for (int jj = 0; jj < IPT; ++jj) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int pi = jj * 8 + s2 * 4 + (lane >> 3), hp = pb * PBLK + pi;
        fgeo[jj][s2] = pi < PBLK ? ((hp / CH_HW) | ((hp % CH_HW) << 8)) : -1;
      }
      const int pi = jj * 8 + (lane & 7), hp = pb * PBLK + pi;
      cgeo[jj] = pi < PBLK ? ((hp / CH_HW) | ((hp % CH_HW) << 8)) : -1;
      coff[jj] = hp * 16;
    }
    const int cch = g * 4 + (lane >> 3);                  // 8-channel chunk (0..7) of the slice this lane transforms
    const uint32_t f_dst0 = stg + (lane >> 3) * 128 + (((lane & 7) ^ (lane >> 3)) * 16);            // instruction 0: pixel l >> 3
    const uint32_t f_dst1 = stg + (4 + (lane >> 3)) * 128 + (((lane & 7) ^ (4 + (lane >> 3))) * 16);  // instruction 1: pixel 4 + (l >> 3)
    const uint32_t c_src0 = stg + (lane & 7) * 128 + (((2 * (lane >> 3)) ^ (lane & 7)) * 16);       // pieces 2c, 2c+1 of pixel l & 7
    const uint32_t c_src1 = stg + (lane & 7) * 128 + (((2 * (lane >> 3) + 1) ^ (lane & 7)) * 16);
    const uint32_t item_off = cch * CH_CHUNK_STRIDE;      // + coff: this lane's 16-byte cell in a 2-byte plane
    const uint32_t item_off8 = (cch >> 1) * CH_CHUNK_STRIDE + (cch & 1) * 8;   // 8-byte cell in a 1-byte plane
    const float* xw = p.x + g * 32 + (lane & 7) * 4;

    // issue the fetches of round jj of slice (n_, th_, tw_, kc_) into staging slot jj % 3 (one commit group per round, possibly empty)
    auto fetch = [&](int jj, bool live, int n_, int th_, int tw_, int kc_) {
      if (live && !(p.dbg & 1)) {
        const float* base = xw + (size_t)n_ * Hs * Ws * p.Cin + kc_ * 64;
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int ge = fgeo[jj][s2];
          const int gh = th_ * CH_TH - 1 + (ge & 0xff), gw = tw_ * CH_TW - 1 + (ge >> 8);
          if (ge >= 0 && (unsigned)gh < (unsigned)p.H && (unsigned)gw < (unsigned)p.W) {
            const float* src = base + (size_t)((gh >> ush) * Ws + (gw >> ush)) * p.Cin;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((s2 ? f_dst1 : f_dst0) + (jj % 3) * 1024), "l"(src) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int n = 0, th = 0, tw = 0, nt = 0;
    if (t_begin < t_end) decode(t_begin, n, th, tw, nt);
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) fetch(jj, t_begin < t_end, n, th, tw, 0);
    uint32_t ai = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      int n2 = n, th2 = th, tw2 = tw, nt2 = nt;            // next tile (prefetch target of the last slice)
      const bool more = t + 1 < t_end;
      if (more) decode(t + 1, n2, th2, tw2, nt2);
      for (int kc = 0; kc < KC; ++kc, ++ai) {
        const int as = ai & 1;
        const bool last_kc = (kc + 1 == KC);
        const bool nx_live = !last_kc || more;
        const int nx_n = last_kc ? n2 : n, nx_th = last_kc ? th2 : th, nx_tw = last_kc ? tw2 : tw, nx_kc = last_kc ? 0 : kc + 1;
        // GroupNorm affine (scale, shift) of this lane's 8 channels: one fetch per slice, consumed by 6 items
        float sc[8], sf[8];
        if (p.affine != nullptr) {
          const float4* ap = reinterpret_cast<const float4*>(p.affine + ((size_t)n * p.Cin + kc * 64 + cch * 8) * 2);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 q4 = __ldg(ap + e);
            sc[2 * e] = q4.x; sf[2 * e] = q4.y; sc[2 * e + 1] = q4.z; sf[2 * e + 1] = q4.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { sc[e] = 1.f; sf[e] = 0.f; }
        }
        mbar_wait(&a_empty[as], ((ai >> 1) & 1) ^ 1);
        const uint32_t dst = smem_u32(sA + as * Cfg::A_SLOT_PAD);
#pragma unroll
        for (int jj = 0; jj < IPT; ++jj) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");          // this lane's pieces of round jj have landed ...
          __syncwarp();                                                  // ... and so have the other lanes'
          const int ge = cgeo[jj];
          float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          const int gh = th * CH_TH - 1 + (ge & 0xff), gw = tw * CH_TW - 1 + (ge >> 8);
          const bool inside = ge >= 0 && (unsigned)gh < (unsigned)p.H && (unsigned)gw < (unsigned)p.W;
          if (inside && !(p.dbg & 2)) {                                  // zero padding applies AFTER the transform
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(c_src0 + (jj % 3) * 1024));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(c_src1 + (jj % 3) * 1024));
          }
          __syncwarp();                                                  // every lane has read the slot: it may be refilled
          // refill the slot just consumed with the round 3 ahead: same slice for jj < 3, next slice (possibly of the next tile) otherwise
          if (jj < 3) fetch(jj + 3, true, n, th, tw, kc);
          else fetch(jj - 3, nx_live, nx_n, nx_th, nx_tw, nx_kc);
          if (ge >= 0 && !(p.dbg & 2)) {
            if (inside) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[e], sf[e]);
              if (p.swish) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {                   // x * sigmoid(x) = x / (1 + 2^(-x log2 e)), approx ex2 / rcp (~1e-7 rel.)
                  float ex, rc;
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(f[e] * -1.4426950408889634f));
                  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
                  f[e] *= rc;
                }
              }
            }
            const uint32_t o16 = dst + item_off + coff[jj];
            if (NPASS == 2) {
              // fp16 plane + two e4m3 planes: lo8 = (x - fp16(x)) * 2^13 and x8 = x (satfinite: out-of-range values degrade gracefully)
              uint32_t h16[4];
              uint32_t l8[2] = {0u, 0u}, x8[2] = {0u, 0u};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h2 = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                h16[e] = *reinterpret_cast<const uint32_t*>(&h2);
                const float2 hf = __half22float2(h2);
                const uint32_t lo2 = __nv_cvt_float2_to_fp8x2(make_float2((f[2 * e] - hf.x) * 8192.0f, (f[2 * e + 1] - hf.y) * 8192.0f), __NV_SATFINITE, __NV_E4M3);
                const uint32_t xx2 = __nv_cvt_float2_to_fp8x2(make_float2(f[2 * e], f[2 * e + 1]), __NV_SATFINITE, __NV_E4M3);
                l8[e >> 1] |= lo2 << (16 * (e & 1));
                x8[e >> 1] |= xx2 << (16 * (e & 1));
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o16), "r"(h16[0]), "r"(h16[1]), "r"(h16[2]), "r"(h16[3]) : "memory");
              const uint32_t o8 = dst + CH_A_PLANE + item_off8 + coff[jj];      // 16-channel chunks of 1-byte elements
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(o8), "r"(l8[0]), "r"(l8[1]) : "memory");
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(o8 + CH_A_PLANE / 2), "r"(x8[0]), "r"(x8[1]) : "memory");
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
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o16 + CH_A_PLANE), "r"(ll[0]), "r"(ll[1]), "r"(ll[2]), "r"(ll[3]) : "memory");
            }
          }
        }
        fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {                                     // operand slot filled
          if (leader) mbar_arrive(&a_full[as]);              // straight to the MMA thread's barrier
          else mbar_arrive(&a_fwd[as]);                      // peer CTA: a local barrier, forwarded across the cluster by warp 3
        }
      }
      n = n2; th = th2; tw = tw2; nt = nt2;
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
