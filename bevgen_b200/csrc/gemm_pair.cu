// 2-CTA (cta_group::2) f16f8 GEMM for the large nn.Linear layers of the stage-2 forward (QKV, MLP1, MLP2; mingpt_sparse.py:170-175,
// 231-237): D[M x N] = A[M x K] . W[N x K]^T with the fp32-equivalent product formed as one fp16 MMA + two e4m3 MMAs per k-step.
//
// Why a second GEMM kernel: gemm_tc (one CTA, 128 x 128 tile, 64 KB of operands per 512-cycle stage, 3 stages) buffers only ~0.84 us of
// tensor work in shared memory, less than the ~1 us TMA round trip, so it runs at ~56 % of the tensor peak (profiles/r01b forward launch
// list).  Here a cluster of two CTAs computes a 256 x 256 tile: each CTA stages ITS 128 rows of A and ITS 128 rows of W (64 KB per stage,
// 3 stages) for 1024 cycles of tensor work per stage - twice the buffered time and half the shared-memory operand bytes per FLOP.
// 256 accumulator columns x 2 buffers fill the TMEM, so the three partial products share ONE accumulator: the operands are pre-scaled
// (A16s = fp16(a * 2^6), W16s = fp16(w * S * 2^7)) so that
//     a*w*2^13*S = A16s*W16s + e4m3((a - a16)*2^13) * e4m3(w*S) + e4m3(a) * e4m3((w - w16)*S*2^13)
// and the epilogue multiplies by out_scale = 1 / (2^13 * S).  |a| < 1024 is required (LayerNorm / GELU outputs).
// Operand planes: A16s [M][K] fp16 + Apair [M][2K bytes], W16s [N][K] fp16 + Wpair [N][2K bytes] (per 64-element k chunk: 64 bytes of
// remainders / scaled values, then 64 bytes of values / scaled remainders, as in gemm_tc npass = 2).
// Warp roles: 0 TMA, 1 MMA issue (whole warp, elected lane), 2 TMEM alloc, 4-11 epilogue (bias, exact-erf GELU, residual, fp32 and/or bf16
// hi/lo planes and/or scaled f16f8 planes for the next GEMM).
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {
namespace gpair {

constexpr int GP_BM = 128, GP_BN = 256, GP_BK = 64, GP_STAGES = 3;
constexpr int GP_TILE = 128 * GP_BK * 2;                  // 16 KB: 128 rows x 64 two-byte elements (or x 128 bytes of e4m3 pairs)
constexpr int GP_STAGE = 4 * GP_TILE;                     // A16, Apair, W16 (this CTA's 128 of the 256 columns), Wpair
constexpr int GP_SMEM = GP_STAGES * GP_STAGE + 1024 + 256;
constexpr int GP_THREADS = 384;                          // warps 4-7 drain accumulator columns 0-127, warps 8-11 columns 128-255

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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GP_THREADS, 1) gemm_pair_f16f8_kernel(const __grid_constant__ GemmPairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + GP_STAGES * GP_STAGE);
  uint64_t* full = bars;                       // GP_STAGES (leader's are used)
  uint64_t* empty = bars + GP_STAGES;          // GP_STAGES (multicast commit: both CTAs)
  uint64_t* tfull = bars + 2 * GP_STAGES;      // 2
  uint64_t* tempty = tfull + 2;                // 2 (leader's collect the 16 epilogue warps of the pair)
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int tiles_m = (p.M + 2 * GP_BM - 1) / (2 * GP_BM), tiles_n = (p.N + GP_BN - 1) / GP_BN;
  const int total = tiles_m * tiles_n;
  const int n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;
  const int kblocks = p.K / GP_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmA[1]); tma_prefetch_desc(&p.tmW[0]); tma_prefetch_desc(&p.tmW[1]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GP_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA: this CTA's 128 rows of A and its 128 of the tile's 256 rows of W, credited to the leader's barrier
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = cid; t < total; t += n_clusters) {           // n tile fastest: clusters running side by side share the A rows in L2
        const int tn = t % tiles_n, tm = t / tiles_n;
        const int arow = tm * 2 * GP_BM + (int)rank * GP_BM, wrow = tn * GP_BN + (int)rank * (GP_BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full[stage], 2 * GP_STAGE);
          uint8_t* st = smem + stage * GP_STAGE;
          tma_load_2d_2sm(st, &p.tmA[0], &full[stage], kb * GP_BK, arow);
          tma_load_2d_2sm(st + GP_TILE, &p.tmA[1], &full[stage], kb * GP_BK, arow);
          tma_load_2d_2sm(st + 2 * GP_TILE, &p.tmW[0], &full[stage], kb * GP_BK, wrow);
          tma_load_2d_2sm(st + 3 * GP_TILE, &p.tmW[1], &full[stage], kb * GP_BK, wrow);
          if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issue (leader CTA): whole warp runs the uniform control flow, one elected lane issues
    if (leader) {
      const bool elected = elect_one();
      const uint32_t idesc = make_idesc_bf16(256, GP_BN, 0, 0) & ~((1u << 7) | (1u << 10));     // fp16 / e4m3 operands, fp32 accumulate
      const uint64_t d0 = make_sdesc_sw128(smem_u32(smem), 16, 1024);
      constexpr uint64_t T = GP_TILE >> 4, ST = GP_STAGE >> 4;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = cid; t < total; t += n_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * GP_BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elected) {
            const uint64_t a16 = d0 + (uint64_t)stage * ST, ap = a16 + T, w16 = a16 + 2 * T, wp = a16 + 3 * T;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_2sm(d, a16 + k * 2, w16 + k * 2, idesc, (kb == 0 && k == 0) ? 0u : 1u);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              umma_f8_2sm(d, ap + k * 2, wp + k * 2, idesc, 1u);                 // e4m3((a - a16) 2^13) * e4m3(w S)
              umma_f8_2sm(d, ap + 4 + k * 2, wp + 4 + k * 2, idesc, 1u);         // e4m3(a) * e4m3((w - w16) S 2^13)
            }
            umma_commit_2sm(&empty[stage]);
          }
          if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elected) umma_commit_2sm(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = row of this CTA's 128-row half, 8 chunks of 32 columns
    const int q = warp & 3, chalf = (warp - 4) >> 2;          // TMEM lane quarter = warp % 4; column half
    uint32_t acc = 0, acc_phase = 0;
    for (int t = cid; t < total; t += n_clusters) {
      const int tn = t % tiles_n, tm = t / tiles_n;
      const int row = tm * 2 * GP_BM + (int)rank * GP_BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const int n0 = tn * GP_BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * GP_BN;
#pragma unroll 1
      for (int c = chalf * (GP_BN / 2); c < (chalf + 1) * (GP_BN / 2); c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        if (c + 32 >= (chalf + 1) * (GP_BN / 2)) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&tempty[acc]);
            else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
          }
        }
        const int col0 = n0 + c;
        if (col0 >= p.N || !row_ok) continue;                 // N % 32 == 0: whole chunks
        float v[32];
        {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = p.bias ? __ldg(bp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * j] = fmaf(__uint_as_float(r[4 * j]), p.out_scale, b4.x); v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), p.out_scale, b4.y);
            v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), p.out_scale, b4.z); v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), p.out_scale, b4.w);
          }
        }
        if (p.gelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        const size_t off = (size_t)row * p.N + col0;
        if (p.residual != nullptr) {
          const float4* rp = reinterpret_cast<const float4*>(p.residual + off);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 rv = rp[j]; v[4 * j] += rv.x; v[4 * j + 1] += rv.y; v[4 * j + 2] += rv.z; v[4 * j + 3] += rv.w; }
        }
        if (p.out_f32 != nullptr) {
          float4* op = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (p.out_hi != nullptr) {                             // bf16 hi / lo planes (the attention kernel's qkv operand)
          uint4* hp = reinterpret_cast<uint4*>(p.out_hi + off);
          uint4* lp = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(v[8 * j + 2 * e], h0, l0);
              split_bf16(v[8 * j + 2 * e + 1], h1, l1);
              h[e] = pack_bf16(h0, h1);
              l[e] = pack_bf16(l0, l1);
            }
            hp[j] = make_uint4(h[0], h[1], h[2], h[3]);
            lp[j] = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        if (p.out_f16 != nullptr) {                            // scaled f16f8 planes for the next GEMM of this kind
          uint4* hp = reinterpret_cast<uint4*>(p.out_f16 + off);
          uint8_t* pp = reinterpret_cast<uint8_t*>(p.out_pair) + (size_t)row * 2 * p.N + (col0 >> 6) * 128 + (col0 & 63);
          uint32_t l8[8], x8[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v0 = v[8 * j + 2 * e], v1 = v[8 * j + 2 * e + 1];
              const __half2 h2 = __floats2half2_rn(v0 * 64.0f, v1 * 64.0f);
              h[e] = *reinterpret_cast<const uint32_t*>(&h2);
              const float2 hf = __half22float2(h2);
              const uint32_t lo2 = __nv_cvt_float2_to_fp8x2(make_float2(fmaf(hf.x, -128.0f, v0 * 8192.0f), fmaf(hf.y, -128.0f, v1 * 8192.0f)), __NV_SATFINITE, __NV_E4M3);
              const uint32_t xx2 = __nv_cvt_float2_to_fp8x2(make_float2(v0, v1), __NV_SATFINITE, __NV_E4M3);
              if (e & 1) { l8[2 * j + (e >> 1)] |= lo2 << 16; x8[2 * j + (e >> 1)] |= xx2 << 16; }
              else { l8[2 * j + (e >> 1)] = lo2; x8[2 * j + (e >> 1)] = xx2; }
            }
            hp[j] = make_uint4(h[0], h[1], h[2], h[3]);
          }
          reinterpret_cast<uint4*>(pp)[0] = make_uint4(l8[0], l8[1], l8[2], l8[3]);
          reinterpret_cast<uint4*>(pp)[1] = make_uint4(l8[4], l8[5], l8[6], l8[7]);
          reinterpret_cast<uint4*>(pp + 64)[0] = make_uint4(x8[0], x8[1], x8[2], x8[3]);
          reinterpret_cast<uint4*>(pp + 64)[1] = make_uint4(x8[4], x8[5], x8[6], x8[7]);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace gpair

int launch_gemm_pair_f16f8(const GemmPairParams& p, int sm_count, cudaStream_t st) {
  if (p.M < 1 || p.N < 32 || (p.N & 31) || p.K < 64 || (p.K & 63)) return BEVGEN_ERR_ARG;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gpair::gemm_pair_f16f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gpair::GP_SMEM) != cudaSuccess) return BEVGEN_ERR_CUDA;
    configured = true;
  }
  const int total = ((p.M + 255) / 256) * ((p.N + 255) / 256);
  const int max_clusters = sm_count / 2;
  const int grid = 2 * (total < max_clusters ? total : max_clusters);
  gpair::gemm_pair_f16f8_kernel<<<grid, gpair::GP_THREADS, gpair::GP_SMEM, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
