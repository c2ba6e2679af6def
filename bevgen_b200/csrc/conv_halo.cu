// 3x3 stride-1 "same" convolution over NHWC bf16 operand planes as a tcgen05 implicit GEMM with a shared-memory HALO tile.
// gemm_tc.cu loads one shifted activation box per tap (9 boxes per 64-channel chunk), which makes shared-memory bandwidth the
// limiter of the bf16x3 product (profiles/r01: tensor pipe 65 %: per k-step the MMAs read 96 KB while TMA writes 64 KB, at 128 B/clk).
// Here the activation halo (16+2) x (8+2) pixels x 64 channels is loaded ONCE per channel chunk into a no-swizzle, 16-byte-chunk-major
// layout [8-channel chunk][halo pixel][16 B]; every tap is then just a different start address of the UMMA A descriptor
// (K-major, SWIZZLE_NONE: core matrices = 8 pixels x 8 channels, LBO = chunk stride, SBO = halo row stride), so activation smem
// write traffic drops 9x (6.4x counting the halo) and the tensor pipe becomes the limiter.  Weights stream through a 4-stage SWIZZLE_128B
// ring exactly as in gemm_tc.cu.  The epilogue fuses bias, residual and the GroupNorm(32) statistics of the OUTPUT (sum / sum of squares
// per (image, group), accumulated per CTA in shared memory and flushed with fp64 atomics when the image changes).
// Output tile = 16 rows x 8 columns of pixels x 128 output channels; CTAs own contiguous tile ranges (halo reuse in L2, few flushes).
#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

constexpr int CH_TW = 8, CH_TH = 16;                       // output tile (pixels)
constexpr int CH_HW = CH_TW + 2, CH_HH = CH_TH + 2;        // halo
constexpr int CH_HPIX = CH_HW * CH_HH;                     // 180 halo pixels
constexpr int CH_CHUNK_STRIDE = CH_HPIX * 16;              // bytes between 8-channel chunks (LBO)
constexpr int CH_ROW_STRIDE = CH_HW * 16;                  // bytes between halo rows (SBO: next 8-pixel core-matrix group)
constexpr int CH_A_PLANE = 8 * CH_CHUNK_STRIDE;            // 23040 B: 64 channels
constexpr int CH_BN = 128;
constexpr int CH_W_TILE = CH_BN * 64 * 2;                  // 16 KB
constexpr int CH_THREADS = 256;

template <int NPASS>
struct ConvHaloCfg {
  static constexpr int NOPS = (NPASS == 3) ? 2 : 1;
  static constexpr int A_SLOT = NOPS * CH_A_PLANE;                       // 46080 / 23040
  static constexpr int A_SLOT_PAD = (A_SLOT + 1023) & ~1023;
  static constexpr int W_STAGE = NOPS * CH_W_TILE;                       // 32 KB / 16 KB
  static constexpr int W_STAGES = (NPASS == 3) ? 3 : 6;
  static constexpr int STATS_BYTES = 4 * 64 * 8 + 4 * 16 * 33 * 4;       // per-warp fp64 accumulators + transpose scratch
  static constexpr int SMEM = 2 * A_SLOT_PAD + W_STAGES * W_STAGE + 1024 + 512 + STATS_BYTES;
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

template <int NPASS>
__global__ void __launch_bounds__(CH_THREADS, 1) conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  using Cfg = ConvHaloCfg<NPASS>;
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (p.W + CH_TW - 1) / CH_TW, tiles_h = (p.H + CH_TH - 1) / CH_TH;
  const int n_tiles_n = (p.Cout + CH_BN - 1) / CH_BN;
  const int KC = p.Cin / 64;
  const long long total = (long long)p.N * tiles_w * tiles_h * n_tiles_n;
  // contiguous tile range per CTA
  const long long t_begin = total * blockIdx.x / gridDim.x, t_end = total * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmW[0]);
    if (NPASS == 3) { tma_prefetch_desc(&p.tmA[1]); tma_prefetch_desc(&p.tmW[1]); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    for (int i = 0; i < WS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  if (threadIdx.x >= 128) { gsm[threadIdx.x - 128] = 0.0; gsm[threadIdx.x] = 0.0; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](long long t, int& n, int& th, int& tw, int& nt) {
    nt = (int)(t % n_tiles_n);
    long long r = t / n_tiles_n;
    tw = (int)(r % tiles_w); r /= tiles_w;
    th = (int)(r % tiles_h);
    n = (int)(r / tiles_h);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ai = 0, wi = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        int n, th, tw, nt;
        decode(t, n, th, tw, nt);
        for (int kc = 0; kc < KC; ++kc, ++ai) {
          const int as = ai & 1;
          mbar_wait(&a_empty[as], ((ai >> 1) & 1) ^ 1);
          mbar_expect_tx(&a_full[as], Cfg::A_SLOT);
#pragma unroll
          for (int o = 0; o < NOPS; ++o)
            tma_load_5d(sA + as * Cfg::A_SLOT_PAD + o * CH_A_PLANE, &p.tmA[o], &a_full[as], 0, tw * CH_TW - 1, th * CH_TH - 1, kc * 8, n);
          for (int tap = 0; tap < 9; ++tap, ++wi) {
            const int ws = wi % WS;
            mbar_wait(&w_empty[ws], ((wi / WS) & 1) ^ 1);
            mbar_expect_tx(&w_full[ws], Cfg::W_STAGE);
#pragma unroll
            for (int o = 0; o < NOPS; ++o)
              tma_load_2d(sW + ws * Cfg::W_STAGE + o * CH_W_TILE, &p.tmW[o], &w_full[ws], kc * 64, tap * p.Cout + nt * CH_BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, CH_BN, 0, 0);
      uint32_t ai = 0, wi = 0, acc = 0, acc_phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * CH_BN;
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
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_hi = make_sdesc_noswz(a_tap + k * 2 * CH_CHUNK_STRIDE, CH_CHUNK_STRIDE, CH_ROW_STRIDE);
              const uint64_t b_hi = make_sdesc_sw128(b_base + k * 32, 16, 1024);
              const uint32_t first = (kc == 0 && tap == 0 && k == 0) ? 0u : 1u;
              if (NPASS == 3) {
                const uint64_t a_lo = make_sdesc_noswz(a_tap + CH_A_PLANE + k * 2 * CH_CHUNK_STRIDE, CH_CHUNK_STRIDE, CH_ROW_STRIDE);
                const uint64_t b_lo = make_sdesc_sw128(b_base + CH_W_TILE + k * 32, 16, 1024);
                umma_bf16(d_tmem, a_lo, b_hi, idesc, first);
                umma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
                umma_bf16(d_tmem, a_hi, b_hi, idesc, 1u);
              } else {
                umma_bf16(d_tmem, a_hi, b_hi, idesc, first);
              }
            }
            umma_commit(&w_empty[ws]);
          }
          umma_commit(&a_empty[as]);
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
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
      const bool row_ok = (ow < p.W) && (oh < p.H);
      const long long roff = (((long long)n * p.H + oh) * p.W + ow) * p.Cout + n0;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * CH_BN;
#pragma unroll 1
      for (int c = 0; c < CH_BN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        if (c + 32 >= CH_BN) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        const int col0 = n0 + c;
        if (col0 >= p.Cout) continue;                   // warp-uniform
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
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

template <int NPASS>
static int launch_conv_halo_t(const ConvHaloParams& p, int sm_count, cudaStream_t st) {
  using Cfg = ConvHaloCfg<NPASS>;
  auto kern = conv_halo_kernel<NPASS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return BEVGEN_ERR_CUDA;
    configured = true;
  }
  const long long total = (long long)p.N * ((p.W + CH_TW - 1) / CH_TW) * ((p.H + CH_TH - 1) / CH_TH) * ((p.Cout + CH_BN - 1) / CH_BN);
  const int grid = (int)(total < sm_count ? total : sm_count);
  kern<<<grid, CH_THREADS, Cfg::SMEM, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_conv_halo(const ConvHaloParams& p, int npass, int sm_count, cudaStream_t st) {
  if (p.Cin % 64 != 0 || p.Cout % 32 != 0 || p.Cout % 4 != 0) return BEVGEN_ERR_ARG;
  if (p.gn_sums != nullptr && (p.Cout / 32 < 4 || p.Cout / 32 > 32 || 32 % (p.Cout / 32) != 0 || p.Cout > 1024)) return BEVGEN_ERR_ARG;
  return npass == 3 ? launch_conv_halo_t<3>(p, sm_count, st) : launch_conv_halo_t<1>(p, sm_count, st);
}

}  // namespace bevgen
