// Persistent KV-cache decode kernel: ONE launch runs every remaining token step of Net2NetTransformer.sample
// (reference modules/stage2/cond_transformer_multi_view.py:154-227; cached formulation of SURVEY.md §3.4) for up to 16 scenes:
//   per step:  24 x { LN1 -> QKV linear -> single-row attention over the KV cache (+ append) -> LN2 -> MLP1 + GELU -> MLP2 + residual }
//              -> LN_f -> head -> top-k / softmax / multinomial -> embedding of the drawn token -> next step.
// One CTA per SM (148), 16 consumer warps + 1 producer thread.  Per step every weight and the whole KV cache cross HBM once, so the
// data side is a BYTE STREAM per SM: every CTA owns a fixed, contiguous slab of each weight matrix (8-row units, packed CTA-major in mma
// fragment order by bevgen_pack_decode_linear) and whole (scene, head) pairs of the KV cache; the producer walks that fixed sequence
// with cp.async.bulk into a 5 x 32 KB shared-memory ring and runs AHEAD of the phases (the bytes do not depend on the
// activations).  The 96 layer phases of a step are NOT separated by grid barriers: everything that crosses CTAs is self-validating
// ("tag sync", see the helpers below); grid-wide barriers (monotonic counter, release at gpu scope) only bracket the head / sampling /
// embedding of a step.  What the kernel's time depends on is the LATENCY CHAIN of its 98 phases per token, not bandwidth (DESIGN.md
// "What bounds the decode kernel"); the rules that came out of the clock traces (tools/decode_trace.py) and same-box A/B runs
// (tools/build_variant.sh):
//   - ONE warp per CTA (per block group in attention) asks an mbarrier, a barrier tells the others (16 warps on one mbarrier serialise);
//   - as few CTA-wide barriers inside a phase as possible (each one re-exposes the slowest warp's L2 latency): two in a linear phase,
//     none between MLP2's K-quarters, one 128-thread barrier per attention block;
//   - global loads that feed the epilogue are issued at the start of the phase, independent ones together; no gpu-scope atomics inside a phase.
//   * activations (16 x d) travel through L2 ALREADY SPLIT into fp16 hi + lo planes in mma A-fragment order: whoever finalises an
//     output element writes it there, and every warp reads ITS k-group of the vector straight into its A registers (8 coalesced
//     16-byte loads per lane).  LayerNorm is applied LAZILY: gamma is folded into the packed weights, the linear runs on the raw vector,
//     and the epilogue applies  rstd_b * (acc - mean_b * c1_n) + c2_n  with c1_n = sum_k gamma_k W_nk, c2_n = bias_n + sum_k beta_k W_nk;
//     the row statistics come from per-unit partial sums the finalisers leave next to the vector (fixed order -> deterministic).
//   * linears: mma.sync m16n8k16, A = the 16 batch rows (fp16 hi + lo), B = 8 weight rows per unit as fp16 + an e4m3 residual plane
//     -> 3 bytes / weight at fp32-equivalent accuracy: x_hi*w16 + x_lo*w16 + x_hi*w8/S, fp32 accumulate; warp w = k-group w of every unit,
//     one cross-warp reduction per phase.  MLP2 is split over K: a CTA works on one quarter of the hidden vector, the finaliser CTA of a
//     row unit polls the four tagged partial sums (mlp2_finalize).
//   * attention: a CTA owns whole (scene, head) pairs; a staged K^T (64 x 128) / V (128 x 64) fp16 block belongs to a group of four warps
//     (32 keys each, CUDA cores, online softmax per warp), the per-warp partials of a pair are merged once; camera-bias row added BEFORE
//     the 1/sqrt(d_head) scale (sparse_self_attention.py:155-168); the newest key is patched into the staged block.
//   * every reduction has a fixed order -> bit-reproducible.
#include <cstdlib>

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

constexpr int DP_CONSUMERS = 512;
constexpr int DP_THREADS = DP_CONSUMERS + 32;
constexpr int DP_NSLOT = 5;
constexpr int DP_SLOT_BYTES = 32768;
constexpr int DP_KG_BYTES = 1536;                       // one 64-wide k-group of an 8-row unit: 2 x 512 B fp16 + 512 B e4m3
constexpr int DP_MAXU = 4;                              // units per reduction batch (16 warps x 128 partial sums each)
constexpr int DP_ACT_BYTES = 65536;                     // one 16 x 1024 activation vector as fp16 hi + lo fragments
constexpr int DP_MAXL = 2560;
constexpr int DP_MAXBH = 4;                             // (scene, head) pairs per CTA: pairs blockIdx.x, blockIdx.x + grid, ...
constexpr int DP_PART = 68;                             // floats per attention partial: m, l, -, -, o[64]
constexpr int DP_MAXV = 4096;
constexpr int DP_MAXLB = DP_MAXL / 16;                  // layout blocks per row at the smallest block size (16)
// act[] outside the activation fetches (float offsets).  Lower 32 KB: cross-warp reduction scratch of the linears | attention scratch;
// upper part: the camera-bias row (requested right after the QKV phase, i.e. while the lower half is still the QKV reduction scratch)
// and the layout row of every pair.
constexpr int DP_A_QS = 0, DP_A_KN = DP_A_QS + DP_MAXBH * 64, DP_A_VN = DP_A_KN + DP_MAXBH * 64, DP_A_PW = DP_A_VN + DP_MAXBH * 64,
              DP_A_TAB = DP_A_PW + 16 * 128, DP_A_MRG = DP_A_TAB + DP_MAXBH * 16 * DP_PART, DP_A_LOW_END = DP_A_MRG + DP_MAXBH * 3 * 64;
constexpr int DP_A_BIAS = 8192, DP_A_LAY = DP_A_BIAS + DP_MAXL, DP_A_END = DP_A_LAY + DP_MAXBH * DP_MAXLB / 4;
static_assert(DP_A_LOW_END <= DP_A_BIAS && DP_MAXU * 16 * 128 <= DP_A_BIAS, "scratch overlaps the prefetched camera-bias row");
static_assert(DP_A_END * 4 <= DP_ACT_BYTES, "attention scratch does not fit the activation buffer");
static_assert((DP_MAXV + 256) * 4 <= DP_ACT_BYTES, "sampling scratch does not fit");

struct DpSmem {
  uint8_t ring[DP_NSLOT][DP_SLOT_BYTES];
  // activation staging (TMA destination, read once into registers) | GEMM cross-warp reduction | attention scratch | sampling scratch
  alignas(128) uint8_t act[DP_ACT_BYTES];
  float2 rowstat[2][16];                                // (mean, rstd) per batch row: [0] of X (LN1 / ln_f), [1] of X1 (LN2)
  float red16[16];
  alignas(8) uint64_t full[DP_NSLOT];
  alignas(8) uint64_t empty[DP_NSLOT];
  alignas(8) uint64_t act_bar;
  alignas(8) uint64_t bias_bar;
  unsigned int flags[48];
  unsigned int slotcnt[DP_NSLOT];                       // warps done with a slot shared by several consumer warps
  int found;
  int rng[4][2];                                        // this CTA's unit range of the QKV / MLP1 / MLP2-row / head linears (part_range, computed once)
  volatile unsigned int rel[DP_NSLOT];                  // unit number last released from each slot (see ring_wait_prev_released)
  volatile unsigned int att_epoch;                      // attention phases (step-major count) whose appends this CTA's consumers have finished (producer gate)
  unsigned int where[4];                                // step, layer, phase of the consumers (diagnostics)
  unsigned long long prof[12];                          // thread 0: ns per phase body / phase boundary (see GPTSampler.PROFILE_SLOTS)
  unsigned long long fine[20];                          // thread 0: ns in activation fetch | linear ring wait | linear mma + epilogue | attention prologue | attention ring wait (warp 0) | attention units | attention merge | MLP2
  unsigned int* debug;                                  // optional pinned host buffer: filled before a timeout trap
  unsigned long long* trace;                            // optional event trace of thread 0 (one layer): (id << 48 | clock) entries
  int trace_n, trace_on;
};

static_assert(sizeof(DpSmem) <= 232448, "DpSmem exceeds the 227 KB of shared memory a CTA can have on sm_100a");

// ------------------------------------------------------------------------------------------------ small helpers
// Event trace of thread 0: probe id + SM clock, recorded while trace_on (one layer of one CTA), read back by tools/decode_trace.py.
// Compiled in only with -DBEVGEN_DP_TRACE (tools/build_variant.sh trace ...): every probe costs a shared-memory load on the critical path.
#ifdef BEVGEN_DP_TRACE
#define DP_TR(smref, id)                                                                                                   \
  do {                                                                                                                     \
    if (threadIdx.x == 0 && (smref).trace_on && (smref).trace_n < 1000) {                                                  \
      (smref).trace[(smref).trace_n++] = ((unsigned long long)(id) << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); \
    }                                                                                                                      \
  } while (0)
#else
#define DP_TR(smref, id) do { } while (0)
#endif
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void bar_group(int g) { asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory"); }

__device__ __forceinline__ void dp_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ unsigned long long dp_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long atom_add_acq_rel_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long old;
  asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned int atom_add_acq_rel_u32(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
// Shared-memory loads the compiler may not sink to their first use: a batch of these is issued back to back, so one warp has several
// loads in flight (ptxas otherwise places every LDS right before its consumer and relies on other warps to cover the 29 cycles).
__device__ __forceinline__ uint32_t lds_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ float lds_f32(const void* p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ uint4 lds_u128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Before a timeout trap: leave (code, CTA, step, layer, phase, a, b) in the caller's pinned host buffer so the failure can be located
__device__ __noinline__ void dp_fail(const DpSmem& sm, unsigned int code, unsigned int a, unsigned int b) {
  volatile unsigned int* dbg = sm.debug;
  if (dbg != nullptr && atomicCAS(const_cast<unsigned int*>(sm.debug), 0u, code) == 0u) {
    dbg[1] = blockIdx.x; dbg[2] = sm.where[0]; dbg[3] = sm.where[1]; dbg[4] = sm.where[2]; dbg[5] = a; dbg[6] = b; dbg[7] = threadIdx.x;
    __threadfence_system();
  }
  __trap();
}

// Grid-wide barrier over the consumer threads of all CTAs (all CTAs are co-resident: one per SM, cooperative launch).  `target` is
// the running arrival count this CTA expects; the counter only grows (zeroed by the host before the launch).  A protocol bug or a
// lost CTA becomes a trap after ~4 s instead of a hung GPU.
__device__ __forceinline__ void grid_sync(DpSmem& sm, unsigned int* counter, unsigned int& target, unsigned int nctas) {
  DP_TR(sm, 90);
  fence_proxy_async();                      // this thread's generic-proxy writes of act[] (scratch) before a later async-proxy refill
  bar_consumers();
  DP_TR(sm, 91);
  target += nctas;
  if (threadIdx.x == 0) {
    red_release_gpu_add(counter, 1u);
    DP_TR(sm, 92);
    const unsigned long long t0 = dp_globaltimer();
    unsigned int polls = 0, seen;
    while ((seen = ld_relaxed_gpu(counter)) < target) {      // no acquire fence (it would invalidate L1, i.e. every spilled register): all cross-CTA data is read with .cg / TMA
      if ((++polls & 1023u) == 0 && dp_globaltimer() - t0 > 4000000000ull) dp_fail(sm, 1u, seen, target);
    }
    DP_TR(sm, 93);
  }
  bar_consumers();
  DP_TR(sm, 94);
}

// Boundary between two layer phases of ONE CTA (no grid barrier: the data are self-validating, see "tag sync" below): the shared scratch
// of the finished phase is free, and its generic-proxy writes of act[] are ordered before a later async-proxy refill.
__device__ __forceinline__ void phase_sync(DpSmem& sm) {
  DP_TR(sm, 90);
  fence_proxy_async();
  bar_consumers();
  DP_TR(sm, 94);
}

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t e4m3x2_to_f16x2(uint16_t v) {
  const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)v, __NV_E4M3);
  return (uint32_t)h.x | ((uint32_t)h.y << 16);
}
// ---- self-validating data ("tag sync") --------------------------------------------------------------------------------------------
// Inside a token step the layer phases are NOT separated by grid barriers.  Every value that crosses CTAs carries a one-bit generation
// tag in its least significant mantissa bit (fp32: 2^-23 relative, the fp16 hi / lo halves of an activation: lo absorbs the forced bit of
// hi, 2^-20 relative), every buffer exists twice (instance = running layer number & 1) and the tag flips each time an instance is
// rewritten.  A reader simply loads what it needs from L2 (ld.global.cg) and repeats the load while a tag is the old one: the poll IS the
// data fetch, no release fence / atomic / second round trip per phase.  Buffer reuse is safe without a barrier because every phase reads
// what the whole grid produced in the phase before: a CTA that writes generation c + 2 of an instance has (transitively) seen every CTA
// finish its reads of generation c.  Real grid barriers remain around the head / sampling / embedding of a step (3 per token).
// Re-use, buffer by buffer (writer of generation c + 2 => every reader of generation c is done; "=>" follows what a CTA must have read):
//   XF / X / PSX  (MLP2 finalisers -> QKV units, attention merge):  finaliser(c + 2) read P2(c + 2) => MLP2 units(c + 2) read HF(c + 2) =>
//                 MLP1 units(c + 2) read all of X1F(c + 2) => every pair owner finished attention(c + 2), i.e. read its QKV(c + 2) => all QKV
//                 units(c + 2) are written (every head has a pair) => their owners finished QKV(c + 1), the readers of XF(c);
//   QKV           (QKV units -> pair owners):  QKV unit(c + 2) read all of XF(c + 1) => all finalisers(c + 1) done => ... all MLP1 units(c + 1)
//                 read all of X1F(c + 1) => every pair owner finished attention(c + 1), hence attention(c);
//   X1F / X1 / PSX1 (pair owners -> MLP1 units, finalisers):  pair owner at attention(c + 2) read QKV(c + 2) => (as above) all finalisers
//                 (c + 1) and all MLP1 units(c + 1) are done, hence MLP1(c) and finalise(c);
//   HF            (MLP1 units -> MLP2 units):  MLP1 unit(c + 2) read all of X1F(c + 2) => pair owners read QKV(c + 2) => QKV units read XF(c + 1)
//                 => finalisers(c + 1) read P2(c + 1) => every MLP2 unit owner finished MLP2(c + 1), hence MLP2(c);
//   P2            (MLP2 units -> finalisers):  MLP2 unit(c + 2) read HF(c + 2) => ... => QKV units read all of XF(c + 1) => all finalisers
//                 (c + 1), hence finalise(c).
// A CTA that owns nothing in a phase reads nothing in it (linear_phase returns before its loads), so it cannot lag behind as a reader.
__device__ __forceinline__ int dp_inst(int c) { return c & 1; }
__device__ __forceinline__ unsigned int dp_tag(int c) { return (unsigned int)(((c >> 1) & 1) ^ 1); }
__device__ __forceinline__ float tag_f32(float v, unsigned int tag) { return __uint_as_float((__float_as_uint(v) & ~1u) | tag); }
__device__ __forceinline__ float ldcg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float2 ldcg_f32x2(const float2* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ldcg_u128(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// bounded polling: a protocol bug becomes a trap after ~4 s instead of a hung GPU
struct DpPoll {
  unsigned int spins = 0;
  unsigned long long t0 = 0ull;
};
__device__ __forceinline__ void poll_tick(const DpSmem& sm, DpPoll& pg, unsigned int code, unsigned int a) {
  if ((++pg.spins & 255u) == 0) {
    const unsigned long long now = dp_globaltimer();
    if (pg.t0 == 0ull) pg.t0 = now;
    else if (now - pg.t0 > 4000000000ull) dp_fail(sm, code, a, pg.spins);
  }
}
// one fp32 value of generation `tag`
__device__ __forceinline__ float ld_tagged(const DpSmem& sm, const float* p, unsigned int tag, unsigned int code) {
  float v = ldcg_f32(p);
  if ((__float_as_uint(v) & 1u) != tag) {
    DpPoll pg;
    do { poll_tick(sm, pg, code, tag); v = ldcg_f32(p); } while ((__float_as_uint(v) & 1u) != tag);
  }
  return v;
}

// x -> fp16 hi, lo with x ~= hi + lo (22 significant bits)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
  const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
  hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
  lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}

// The consumer phases are inlined into one loop nest; without this the compiler hoists every phase's thread- and size-derived constants
// out of the step / layer loops and spills them.  Values that pass through an (empty) volatile asm are recomputed where they are used.
__device__ __forceinline__ int opaque(int v) { asm volatile("" : "+r"(v)); return v; }

__device__ __forceinline__ void part_range(int U, int& u0, int& u1) {
  const int bx = opaque((int)blockIdx.x), g = opaque((int)gridDim.x);
  u0 = (int)(((long long)bx * U) / g);
  u1 = (int)(((long long)(bx + 1) * U) / g);
}
// MLP2: quarter q of the K = 4d hidden vector belongs to CTAs [q * (grid / 4), (q + 1) * (grid / 4)), which share its d / 8 row units
__device__ __forceinline__ void quarter_range(int upq, int& u0, int& u1) {
  const int cpq = (int)gridDim.x >> 2, q = (int)blockIdx.x / cpq, j = (int)blockIdx.x - q * cpq;
  if (q >= 4) { u0 = u1 = 0; return; }
  u0 = q * upq + (int)(((long long)j * upq) / cpq);
  u1 = q * upq + (int)(((long long)(j + 1) * upq) / cpq);
}
enum { RNG_QKV = 0, RNG_MLP1 = 1, RNG_MLP2 = 2, RNG_HEAD = 3 };
__device__ __forceinline__ void cta_range(const DpSmem& sm, int which, int& u0, int& u1) { u0 = sm.rng[which][0]; u1 = sm.rng[which][1]; }
__device__ __forceinline__ int cta_of_unit(int u, int U) { return (int)((((long long)(u + 1)) * gridDim.x - 1) / U); }

__device__ __forceinline__ float consumers_sum(float v, float* red16) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  bar_consumers();
  if ((threadIdx.x & 31) == 0) red16[threadIdx.x >> 5] = v;
  bar_consumers();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += red16[i];
  return t;
}
__device__ __forceinline__ float consumers_max(float v, float* red16) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  bar_consumers();
  if ((threadIdx.x & 31) == 0) red16[threadIdx.x >> 5] = v;
  bar_consumers();
  float t = red16[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) t = fmaxf(t, red16[i]);
  return t;
}

// mbarrier wait with a wall-clock bound: a ring-protocol bug traps after ~4 s instead of hanging the GPU
__device__ __forceinline__ void dp_mbar_wait(DpSmem& sm, uint64_t* bar, uint32_t parity, unsigned int code, unsigned int seq) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = dp_globaltimer();
  unsigned int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0 && dp_globaltimer() - t0 > 4000000000ull) dp_fail(sm, code, seq, parity);
  }
}
// ring bookkeeping shared by the producer and the consumers: unit number `seq` lives in slot seq % NSLOT, phase (seq / NSLOT) & 1
__device__ __forceinline__ void ring_wait_full(DpSmem& sm, unsigned int seq) { dp_mbar_wait(sm, &sm.full[seq % DP_NSLOT], (seq / DP_NSLOT) & 1u, 2u, seq); }
__device__ __forceinline__ void ring_release(DpSmem& sm, unsigned int seq) {
  sm.rel[seq % DP_NSLOT] = seq;
  __threadfence_block();
  mbar_arrive(&sm.empty[seq % DP_NSLOT]);
}
// mbarrier parity waits alias modulo 2 rounds.  A single in-order consumer can never be a round early, but the attention phase has four
// warp groups taking units out of order: a group could test full[slot] for round j + 1 while round j has not even landed, and the
// parity test would answer with round j - 1's completion.  So a group first waits until the slot's previous occupant (unit seq - NSLOT)
// has been RELEASED (then round j is over and round j + 2 cannot start before this unit is released): the parity wait is unambiguous.
__device__ __forceinline__ void ring_wait_prev_released(DpSmem& sm, unsigned int seq) {
  if (seq < (unsigned)DP_NSLOT) return;
  const unsigned int want = seq - DP_NSLOT, slot = seq % DP_NSLOT;
  if (sm.rel[slot] == want) return;
  const unsigned long long t0 = dp_globaltimer();
  unsigned int spins = 0;
  while (sm.rel[slot] != want) {
    if ((++spins & 4095u) == 0 && dp_globaltimer() - t0 > 4000000000ull) dp_fail(sm, 4u, seq, sm.rel[slot]);
  }
}

// A slot consumed by `nwarps` independent warps: each calls this from lane 0 after its last read; the last one hands the slot back.
__device__ __forceinline__ void ring_release_shared(DpSmem& sm, unsigned int seq, unsigned int nwarps) {
  __threadfence_block();
  const unsigned int slot = seq % DP_NSLOT;
  if (nwarps == 1u) { ring_release(sm, seq); return; }
  if (atomicAdd(&sm.slotcnt[slot], 1u) == nwarps - 1u) {
    sm.slotcnt[slot] = 0u;
    ring_release(sm, seq);
  }
}

// ------------------------------------------------------------------------------------------------ producer
// The byte stream of this CTA, in exactly the order its consumers take it (both sides derive it from blockIdx alone):
//   per layer: QKV units [part_range(3d/8)] | every 128-key block of the CTA's (scene, head) pairs | MLP1 units [part_range(4d/8)] |
//              MLP2 units of the CTA's K-quarter [quarter_range];   per step: head units [part_range(vpad/8)].
__device__ __noinline__ void dp_producer(const DecodeParams& p, DpSmem& sm) {
  unsigned int seq = 0;
  const int d = p.d, KG = d >> 6;
  const uint32_t unit_bytes = (uint32_t)KG * DP_KG_BYTES;
  unsigned long long tw = 0ull, ta = 0ull, te = 0ull;
  const bool prof = p.profile != nullptr;
  auto issue_unit = [&](const uint8_t* src) {
    const unsigned int slot = seq % DP_NSLOT;
    if (seq >= (unsigned)DP_NSLOT) {
      const unsigned long long t0 = prof ? dp_globaltimer() : 0ull;
      dp_mbar_wait(sm, &sm.empty[slot], ((seq / DP_NSLOT) - 1u) & 1u, 3u, seq);
      if (prof) tw += dp_globaltimer() - t0;
    }
    if (p.dbg & 8) mbar_arrive(&sm.full[slot]);
    else {
      mbar_expect_tx(&sm.full[slot], unit_bytes);
      dp_bulk_g2s(sm.ring[slot], src, unit_bytes, &sm.full[slot]);
    }
    ++seq;
  };
  auto stream_units = [&](const uint8_t* base, int which) {
    int u0, u1;
    cta_range(sm, which, u0, u1);
    for (int u = u0; u < u1; ++u) issue_unit(base + (size_t)u * unit_bytes);
  };
  const int BH = p.B * p.H;
  for (int s = p.step_begin; s < p.step_end; ++s) {
    const int n = p.nc + s, nblk = (n + 127) >> 7;
    for (int l = 0; l < p.n_layers; ++l) {
      const DecodeLayer& L = p.layers[l];
      stream_units(L.w_qkv, RNG_QKV);
      if ((int)blockIdx.x < BH) {  // attention units: every 128-key block of this CTA's (scene, head) pairs
        if (s > p.step_begin) {
          // The blocks hold keys this CTA's consumers appended during step s - 1: do not run ahead of their attention phase (s - 1, l).
          // With 24 layers the ring (5 units) never reaches that far back; tiny models (a few units per step) do.
          const unsigned int need = (unsigned)(s - 1 - p.step_begin) * (unsigned)p.n_layers + (unsigned)l + 1u;
          if (sm.att_epoch < need) {
            const unsigned long long t0 = dp_globaltimer();
            const unsigned long long te0 = t0;
            unsigned int spins = 0;
            while (sm.att_epoch < need) {
              if ((++spins & 4095u) == 0 && dp_globaltimer() - t0 > 4000000000ull) dp_fail(sm, 5u, need, sm.att_epoch);
            }
            te += dp_globaltimer() - te0;
          }
          __threadfence();                  // the consumers' appends (same CTA; observed through att_epoch) are performed at gpu scope ...
          fence_proxy_async_all();          // ... and ordered before this thread's async-proxy reads
        }
        for (int bh = blockIdx.x; bh < BH; bh += gridDim.x) {
          for (int blk = 0; blk < nblk; ++blk, ++seq) {
            const unsigned int slot = seq % DP_NSLOT;
            if (seq >= (unsigned)DP_NSLOT) {
              const unsigned long long t0 = prof ? dp_globaltimer() : 0ull;
              dp_mbar_wait(sm, &sm.empty[slot], ((seq / DP_NSLOT) - 1u) & 1u, 3u, seq);
              if (prof) ta += dp_globaltimer() - t0;
            }
            const int cnt = min(128, n - blk * 128);
            const uint32_t kbytes = 64u * 128u * 2u, vbytes = (uint32_t)cnt * 128u;
            if (p.dbg & 8) { mbar_arrive(&sm.full[slot]); continue; }
            mbar_expect_tx(&sm.full[slot], kbytes + vbytes);
            const __half* kc = reinterpret_cast<const __half*>(L.kc) + ((size_t)bh * (p.Lmax >> 7) + blk) * (64 * 128);
            const __half* vc = reinterpret_cast<const __half*>(L.vc) + ((size_t)bh * p.Lmax + (size_t)blk * 128) * 64;
            dp_bulk_g2s(sm.ring[slot], kc, kbytes, &sm.full[slot]);
            dp_bulk_g2s(sm.ring[slot] + kbytes, vc, vbytes, &sm.full[slot]);
          }
        }
      }
      stream_units(L.w_1, RNG_MLP1);
      stream_units(L.w_2, RNG_MLP2);        // the CTA's row units of ITS K-quarter (unit number = quarter * d / 8 + row unit)
    }
    stream_units(p.w_head, RNG_HEAD);
  }
  if (p.profile != nullptr) {            // producer: ns blocked on a full ring (weight units | attention units), on the epoch gate
    p.profile[(size_t)blockIdx.x * 32 + 22] = tw; p.profile[(size_t)blockIdx.x * 32 + 23] = ta; p.profile[(size_t)blockIdx.x * 32 + 24] = te;
  }
}

// ------------------------------------------------------------------------------------------------ activation vectors in fragment order
// A [16 x d] vector lives in global memory as d/64 k-groups of 4 KB: [k-group][j = 0..3: fp16 hi of k-step j | 4..7: fp16 lo][lane][16 B],
// the 16 bytes of a lane being its four mma A registers (a0: row g, cols 2t,2t+1 | a1: row g+8 | a2: row g, cols 2t+8,2t+9 | a3: row g+8).
__device__ __forceinline__ int frag_off(int b, int c) {
  const int kg = c >> 6, ks = (c >> 4) & 3, cc = c & 15;
  const int lane = ((b & 7) << 2) | ((cc & 7) >> 1), reg = (b >> 3) | ((cc >> 3) << 1);
  return kg * 4096 + ks * 512 + lane * 16 + reg * 4 + (cc & 1) * 2;
}
// Both halves carry the generation tag in their last mantissa bit; lo is computed against the TAGGED hi, so hi + lo still has ~20 bits.
__device__ __forceinline__ void store_frag(uint8_t* __restrict__ base, int b, int c, float x, unsigned int tag) {
  const unsigned short hb = (unsigned short)((__half_as_ushort(__float2half_rn(x)) & 0xfffeu) | tag);
  const unsigned short lb = (unsigned short)((__half_as_ushort(__float2half_rn(x - __half2float(__ushort_as_half(hb)))) & 0xfffeu) | tag);
  const int off = frag_off(b, c);
  *reinterpret_cast<unsigned short*>(base + off) = hb;
  *reinterpret_cast<unsigned short*>(base + off + 2048) = lb;
}
// ------------------------------------------------------------------------------------------------ linear phase (consumers)
enum { EPI_QKV = 0, EPI_MLP1 = 1, EPI_M2P = 2, EPI_HEAD = 3 };      // EPI_M2P: a K-quarter of MLP2 -> raw partial sums (no LayerNorm, no bias)

// Inputs of a linear phase in ONE round trip: the warp's k-group of the activation vector (fragment order, 8 coalesced 16-byte loads per lane
// straight into the A registers; every half of the rows that exist must carry `tag`) and the LayerNorm partial sums ps[row][part][2] of
// batch row w (warp w = row w; nparts <= 128: four loads per lane).  All twelve loads are issued before the first tag is looked at,
// whatever is still of the old generation is asked for again.
__device__ __forceinline__ void load_inputs_tagged(const DecodeParams& p, const DpSmem& sm, const float* __restrict__ ps, int nparts, bool do_stats,
                                                   const uint8_t* __restrict__ vec, bool do_frag, int w, int lane, unsigned int tag, uint32_t (&ahi)[4][4],
                                                   uint32_t (&alo)[4][4], float& S, float& Q) {
  const uint4* src = reinterpret_cast<const uint4*>(vec + (size_t)w * 4096 + lane * 16);
  const uint32_t tm = tag ? 0x00010001u : 0u;
  const int g = lane >> 2;
  const uint32_t m0 = (g < p.B) ? 0x00010001u : 0u, m1 = (g + 8 < p.B) ? 0x00010001u : 0u;      // registers 0, 2: row g | 1, 3: row g + 8
  bool need_s = do_stats && w < p.B, need_f = do_frag;
  float2 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = make_float2(0.f, 0.f);
  DpPoll pg;
  for (;;) {
    if (need_s) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = lane + 32 * j;
        if (i < nparts) v[j] = ldcg_f32x2(reinterpret_cast<const float2*>(ps) + (size_t)w * nparts + i);
      }
    }
    uint32_t bad_f = 0u;
    if (need_f) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 h4 = ldcg_u128(src + j * 32), l4 = ldcg_u128(src + 128 + j * 32);
        ahi[j][0] = h4.x; ahi[j][1] = h4.y; ahi[j][2] = h4.z; ahi[j][3] = h4.w;
        alo[j][0] = l4.x; alo[j][1] = l4.y; alo[j][2] = l4.z; alo[j][3] = l4.w;
        bad_f |= (((h4.x ^ tm) | (h4.z ^ tm) | (l4.x ^ tm) | (l4.z ^ tm)) & m0) | (((h4.y ^ tm) | (h4.w ^ tm) | (l4.y ^ tm) | (l4.w ^ tm)) & m1);
      }
    }
    unsigned int bad_s = 0u;
    if (need_s) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane + 32 * j < nparts) bad_s |= ((__float_as_uint(v[j].x) ^ tag) | (__float_as_uint(v[j].y) ^ tag)) & 1u;
    }
    need_s = need_s && bad_s != 0u;
    need_f = need_f && bad_f != 0u;
    if (!need_s && !need_f) break;
    poll_tick(sm, pg, 8u, (unsigned)w);
  }
  __syncwarp();
  S = (v[0].x + v[1].x) + (v[2].x + v[3].x);
  Q = (v[0].y + v[1].y) + (v[2].y + v[3].y);
}

// LayerNorm statistics of the 16 rows from the finalisers' partial sums ps[row][part][2] = (sum, sum of squares): warp w = row w.
// Two halves so that the global loads (issued at the start of a phase) are in flight while the activation vector arrives and the MMAs run.
__device__ __forceinline__ void row_stats_finish(const DecodeParams& p, DpSmem& sm, int which, float S, float Q) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = 16; o; o >>= 1) { S += __shfl_xor_sync(0xffffffffu, S, o); Q += __shfl_xor_sync(0xffffffffu, Q, o); }
  if (lane == 0) {
    const float inv_d = 1.0f / (float)p.d, mean = S * inv_d;
    sm.rowstat[which][w] = make_float2(mean, rsqrtf(fmaxf(Q * inv_d - mean * mean, 0.f) + 1e-5f));
  }
}

// The CTA's copy of an activation vector (fragment order, in act[]) -> the A registers of warp w's 64-wide k-group
__device__ __forceinline__ void load_afrag(const DpSmem& sm, int w, int lane, uint32_t (&ahi)[4][4], uint32_t (&alo)[4][4]) {
  const uint8_t* base = sm.act + w * 4096 + lane * 16;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 h4 = *reinterpret_cast<const uint4*>(base + j * 512), l4 = *reinterpret_cast<const uint4*>(base + 2048 + j * 512);
    ahi[j][0] = h4.x; ahi[j][1] = h4.y; ahi[j][2] = h4.z; ahi[j][3] = h4.w;
    alo[j][0] = l4.x; alo[j][1] = l4.y; alo[j][2] = l4.z; alo[j][3] = l4.w;
  }
}
// Warp w's k-group of one staged 8-row unit: three independent accumulator chains x_hi * w16, x_hi * w8 (scaled by 1 / S afterwards), x_lo * w16
__device__ __forceinline__ void unit_mma(const uint8_t* slot, int w, int lane, const uint32_t (&ahi)[4][4], const uint32_t (&alo)[4][4], float (&acc0)[4],
                                         float (&acc1)[4], float (&acc2)[4]) {
  const uint8_t* base = slot + w * DP_KG_BYTES + lane * 16;
  const uint4 h0 = *reinterpret_cast<const uint4*>(base), h1 = *reinterpret_cast<const uint4*>(base + 512);
  const uint4 lo = *reinterpret_cast<const uint4*>(base + 1024);
  mma_f16(acc0, ahi[0], h0.x, h0.y); mma_f16(acc2, alo[0], h0.x, h0.y);
  mma_f16(acc1, ahi[0], e4m3x2_to_f16x2((uint16_t)(lo.x & 0xffffu)), e4m3x2_to_f16x2((uint16_t)(lo.x >> 16)));
  mma_f16(acc0, ahi[1], h0.z, h0.w); mma_f16(acc2, alo[1], h0.z, h0.w);
  mma_f16(acc1, ahi[1], e4m3x2_to_f16x2((uint16_t)(lo.y & 0xffffu)), e4m3x2_to_f16x2((uint16_t)(lo.y >> 16)));
  mma_f16(acc0, ahi[2], h1.x, h1.y); mma_f16(acc2, alo[2], h1.x, h1.y);
  mma_f16(acc1, ahi[2], e4m3x2_to_f16x2((uint16_t)(lo.z & 0xffffu)), e4m3x2_to_f16x2((uint16_t)(lo.z >> 16)));
  mma_f16(acc0, ahi[3], h1.z, h1.w); mma_f16(acc2, alo[3], h1.z, h1.w);
  mma_f16(acc1, ahi[3], e4m3x2_to_f16x2((uint16_t)(lo.w & 0xffffu)), e4m3x2_to_f16x2((uint16_t)(lo.w >> 16)));
}

// One linear layer (QKV, MLP1, head) on this CTA's units: acc[b][row] = sum_k act[b][k] * W'[row][k]; `seq` advances by the units consumed.
// Warp w owns the k-group w of EVERY unit (its A fragments stay in registers for the whole phase), takes a unit as soon as it has landed
// and hands the slot back through a shared-memory counter (no CTA barrier per unit); the 16 partial sums of up to DP_MAXU units are added
// once, one output element per thread, whose epilogue constants were requested before the MMAs.
//   frag_src: the activation vector in fragment order;  stats / nparts / which: partial sums of the lazy LayerNorm, see combine_row_stats
//   rtag: generation tag of the inputs (frag_src, stats);  out / wtag: where the results go (QKV: fp32 [16][3d], MLP1: the four fragment-
//   order quarters, head: LOGITS - untagged, a real grid barrier follows) and the tag they carry
__device__ __noinline__ void linear_phase(const DecodeParams& p, DpSmem& sm, unsigned int& seq, unsigned int& act_par, const int epi, const uint8_t* __restrict__ frag_src,
                             const float* __restrict__ stats, int nparts, int which, int rng, float inv_s, const float* __restrict__ c1,
                             const float* __restrict__ c2, unsigned int rtag, void* __restrict__ out, unsigned int wtag) {
  const int tid = opaque((int)threadIdx.x), w = tid >> 5, lane = tid & 31;
  const int d = opaque(p.d), KG = d >> 6;
  const uint32_t vec_bytes = (uint32_t)KG * 4096u;
  float* red = reinterpret_cast<float*>(sm.act);
  DP_TR(sm, 10);
  int u0, u1;
  cta_range(sm, rng, u0, u1);
  DP_TR(sm, 11);
  unsigned long long tf = p.profile != nullptr ? dp_globaltimer() : 0ull;
  auto fine = [&](int k) { if (tid == 0 && p.profile != nullptr && !sm.trace) { const unsigned long long now = dp_globaltimer(); sm.fine[k] += now - tf; tf = now; } };
  float rsS = 0.f, rsQ = 0.f;
  // output element of a unit owned by this thread: element e -> lane e / 4, register e % 4 of the accumulator fragment -> (batch row, weight row);
  // the epilogue constants of the first batch are requested before anything else (consumed after the MMAs)
  const int ui = tid >> 7, e = tid & 127;
  const int ln = e >> 2, j = e & 3;
  const int b = (ln >> 2) + ((j & 2) ? 8 : 0), nrow = 2 * (ln & 3) + (j & 1);
  float c1f = 0.f, c2f = 0.f;
  if (epi != EPI_M2P && u0 + ui < u1 && b < p.B) { c1f = __ldg(c1 + (u0 + ui) * 8 + nrow); c2f = __ldg(c2 + (u0 + ui) * 8 + nrow); }
  DP_TR(sm, 12);
  const bool lnorm = epi != EPI_M2P;
  // (a CTA without units only needs the statistics when its attention phase will: LN1 of the pairs it owns)
  if (u1 == u0 && !(lnorm && which == 0 && (int)blockIdx.x < p.B * p.H)) return;
  // Every warp reads ITS k-group of the activation vector straight from L2 into the A registers (fragment order: 8 coalesced 16-byte
  // loads per lane; .cg: the vector was written by other SMs in the previous phase) - no staging buffer, no barrier for it.  The
  // statistics' partial sums (consumed after the MMAs) and the fragments are requested TOGETHER: one L2 round trip, not two.
  uint32_t ahi[4][4], alo[4][4];
  load_inputs_tagged(p, sm, stats, nparts, lnorm, frag_src, u1 > u0 && w < KG, w, lane, rtag, ahi, alo, rsS, rsQ);
  DP_TR(sm, 13);
  if (u1 == u0) { row_stats_finish(p, sm, which, rsS, rsQ); return; }
  // ONE warp asks each unit's mbarrier (16 warps asking the same barrier serialise in the SM's sync unit: ~1500 cycles per wait in the
  // trace of tools/decode_trace.py), warp k the one of unit k of the batch; the others learn through the CTA barrier.  The units were
  // requested phases ago, but even a wait that returns at once costs ~250 cycles: four in a row by warp 0 were ~1 K cycles in front of
  // the first CTA barrier of every linear phase.
  if (w < min(DP_MAXU, u1 - u0)) ring_wait_full(sm, seq + (unsigned)w);
  DP_TR(sm, 15);
  bar_consumers();                          // the first batch of units has landed; act[] (reduction scratch) is free: the previous phase's readers are behind this or an earlier CTA barrier
  DP_TR(sm, 16);
  for (int ub = u0; ub < u1; ub += DP_MAXU) {
    const int nb = min(DP_MAXU, u1 - ub);
    const bool has = ui < nb && b < p.B;
    const int row = (ub + ui) * 8 + nrow;
    float c1v = c1f, c2v = c2f;
    if (ub != u0 && has && lnorm) { c1v = __ldg(c1 + row); c2v = __ldg(c2 + row); }
    if (ub != u0) {                          // a further batch (more than DP_MAXU units per CTA: small grids only)
      if (w < nb) ring_wait_full(sm, seq + (unsigned)w);
      bar_consumers();
    }
    if (w < KG) {
      for (int k = 0; k < nb; ++k) {
        const unsigned int sq = seq + (unsigned)k;
        DP_TR(sm, 17);
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
        if (!(p.dbg & 2)) unit_mma(sm.ring[sq % DP_NSLOT], w, lane, ahi, alo, acc0, acc1, acc2);
        *reinterpret_cast<float4*>(&red[((k * 16 + w) * 32 + lane) * 4]) =
            make_float4((acc0[0] + acc2[0]) + acc1[0] * inv_s, (acc0[1] + acc2[1]) + acc1[1] * inv_s, (acc0[2] + acc2[2]) + acc1[2] * inv_s,
                        (acc0[3] + acc2[3]) + acc1[3] * inv_s);
        fine(2);
        DP_TR(sm, 18);
      }
    }
    if (ub == u0 && lnorm) row_stats_finish(p, sm, which, rsS, rsQ);      // published by the barrier below
    DP_TR(sm, 19);
    bar_consumers();                        // every warp's partial sums are written, i.e. its loads of the staged units have returned
    if (tid == 0) for (int k = 0; k < nb; ++k) ring_release(sm, seq + (unsigned)k);      // one thread hands the batch's slots back
    seq += (unsigned)nb;
    DP_TR(sm, 20);
    if (has) {
      // (ii) four independent partial sums: the 16 shared-memory loads are in flight together
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      const float* rp = red + (ui * 16) * 128 + e;
      int ww = 0;
      for (; ww + 4 <= KG; ww += 4) { v0 += rp[ww * 128]; v1 += rp[(ww + 1) * 128]; v2 += rp[(ww + 2) * 128]; v3 += rp[(ww + 3) * 128]; }
      for (; ww < KG; ++ww) v0 += rp[ww * 128];
      float v = (v0 + v1) + (v2 + v3);
      if (lnorm) {
        const float2 st = sm.rowstat[which][b];
        v = st.y * (v - st.x * c1v) + c2v;                                     // lazy LayerNorm + bias
      }
      if (epi == EPI_QKV) reinterpret_cast<float*>(out)[(size_t)b * 3 * d + row] = tag_f32(v, wtag);
      else if (epi == EPI_MLP1) { const int qq = row / d; store_frag(reinterpret_cast<uint8_t*>(out) + (size_t)qq * vec_bytes, b, row - qq * d, gelu_erf(v), wtag); }
      else if (epi == EPI_M2P) { const int qq = row / d; reinterpret_cast<float*>(out)[((size_t)qq * 16 + b) * d + (row - qq * d)] = tag_f32(v, wtag); }
      else if (row < p.vocab) p.LOGITS[(size_t)b * p.vocab + row] = v;
    }
    DP_TR(sm, 21);
    if (ub + DP_MAXU < u1) bar_consumers();   // the reduction scratch is reused by the next batch
  }
  fine(2);
}

// MLP2 + residual, split over K: the four K-quarters of the hidden vector belong to four groups of gridDim / 4 CTAs; a CTA runs its row
// units of ITS quarter as an ordinary linear phase (EPI_M2P: one 64 KB activation quarter per CTA instead of the whole 256 KB vector - the
// 148-fold broadcast of the hidden vector was what the phase's time went into) and leaves tagged fp32 partial sums P2[quarter][16][d].
// CTA ru < d / 8 then finishes row unit ru: residual + bias + the four partials in quarter order (polled like any other tagged value: no
// atomic, no fence), the new residual rows, their fragment-order copy and the unit's LayerNorm partial sums.
__device__ __noinline__ void mlp2_finalize(const DecodeParams& p, DpSmem& sm, const DecodeLayer& L, int c) {
  const int tid = opaque((int)threadIdx.x);
  const int d = opaque(p.d), upq = d >> 3, ru = opaque((int)blockIdx.x);
  if (ru >= upq || tid >= 128) return;
  const int inst = dp_inst(c);
  const unsigned int tag = dp_tag(c);
  const float* X1i = p.X1 + (size_t)inst * 16 * d;
  const float* P2i = p.P2 + (size_t)inst * 4 * 16 * d;
  float* Xo = p.X + (size_t)inst * 16 * d;
  uint8_t* XFo = p.XF + (size_t)inst * (size_t)(d >> 6) * 4096;
  float* PSXo = p.PSX + (size_t)inst * 16 * upq * 2;
  const int fb = tid >> 3, fn = tid & 7;                 // 8 consecutive lanes = one batch row
  const int row = ru * 8 + fn;
  const bool fin = fb < p.B;
  float v = 0.f;
  if (fin) {
    const float bias = __ldg(L.c2_2 + row);
    float x1, pq[4];
    DpPoll pg;
    for (;;) {                              // five independent loads in flight; repeated while a tag is old
      x1 = ldcg_f32(X1i + (size_t)fb * d + row);
#pragma unroll
      for (int q = 0; q < 4; ++q) pq[q] = ldcg_f32(P2i + ((size_t)q * 16 + fb) * d + row);
      const unsigned int bad = ((__float_as_uint(x1) ^ tag) | (__float_as_uint(pq[0]) ^ tag) | (__float_as_uint(pq[1]) ^ tag) | (__float_as_uint(pq[2]) ^ tag) |
                                (__float_as_uint(pq[3]) ^ tag)) & 1u;
      if (bad == 0u) break;
      poll_tick(sm, pg, 13u, (unsigned)ru);
    }
    v = (x1 + bias) + ((pq[0] + pq[1]) + (pq[2] + pq[3]));
    Xo[(size_t)fb * d + row] = tag_f32(v, tag);
    store_frag(XFo, fb, row, v, tag);
  }
  __syncwarp();
  float sv = v, qv = v * v;
  sv += __shfl_xor_sync(0xffffffffu, sv, 1); qv += __shfl_xor_sync(0xffffffffu, qv, 1);
  sv += __shfl_xor_sync(0xffffffffu, sv, 2); qv += __shfl_xor_sync(0xffffffffu, qv, 2);
  sv += __shfl_xor_sync(0xffffffffu, sv, 4); qv += __shfl_xor_sync(0xffffffffu, qv, 4);
  if (fin && fn == 0) *reinterpret_cast<float2*>(PSXo + ((size_t)fb * upq + ru) * 2) = make_float2(tag_f32(sv, tag), tag_f32(qv, tag));
}

// ------------------------------------------------------------------------------------------------ attention phase (consumers)
// The finished attention output of a pair gives x1 = LN1(x) + attention (Block.forward takes the residual from the LayerNorm OUTPUT,
// mingpt_sparse.py:242-245) -> fp32 X1 (MLP2's residual), fragment-ordered X1F (MLP1's operand) and the head's partial LayerNorm sums of the row.
// One CTA owns whole (scene, head) pairs (pairs blockIdx.x, blockIdx.x + grid, ...: no partial results cross CTAs).  A staged 128-key block
// belongs to a group of four warps, each warp to 32 of its keys from the moment the block lands until the slot goes back - no barrier inside
// a block: lane = (key pair, channel half) for q . K^T (4-byte loads of the transposed block against q broadcast from shared memory, one
// shuffle joins the halves), half-warp online softmax, lane = (key mod 4, 8 channels) for P . V (16-byte loads of the value rows,
// probabilities through a 128-byte per-warp scratch).  Blocks go round-robin over the four groups, every warp keeps a running
// (max, sum, o[64]) per pair and leaves it in a table that one warp per pair merges in warp order.
__device__ __noinline__ void attention_phase(const DecodeParams& p, DpSmem& sm, unsigned int& seq, unsigned int& bias_par, const DecodeLayer& L, int s, int c,
                                            bool first_layer) {
  const int tid = opaque((int)threadIdx.x), w = tid >> 5, lane = tid & 31;
  const int d = opaque(p.d), H = opaque(p.H), BH = p.B * H, G = opaque((int)gridDim.x), bx = opaque((int)blockIdx.x);
  const int n = p.nc + s, r = n - 1, nblk = (n + 127) >> 7;
  const int npairs = bx < BH ? (BH - 1 - bx) / G + 1 : 0;
  if (npairs == 0) return;
  const int nun = npairs * nblk;
  // inputs: QKV of this layer (generation c), the residual stream X of the layer before (generation c - 1: MLP2 or the embedding)
  const unsigned int tag = dp_tag(c), ptag = dp_tag(c - 1);
  const float* QKVi = p.QKV + (size_t)dp_inst(c) * 16 * 3 * d;
  const float* Xi = p.X + (size_t)dp_inst(c - 1) * 16 * d;
  float* X1o = p.X1 + (size_t)dp_inst(c) * 16 * d;
  uint8_t* X1Fo = p.X1F + (size_t)dp_inst(c) * (size_t)(d >> 6) * 4096;
  float* PSX1o = p.PSX1 + (size_t)dp_inst(c) * 16 * H * 2;
  unsigned long long tf = p.profile != nullptr ? dp_globaltimer() : 0ull;
  auto fine = [&](int k) { if (tid == 0 && p.profile != nullptr && !sm.trace) { const unsigned long long now = dp_globaltimer(); sm.fine[k] += now - tf; tf = now; } };
  float* fa = reinterpret_cast<float*>(sm.act);
  float* qs = fa + DP_A_QS;                            // [MAXBH][64]   q of every pair
  float* kn = fa + DP_A_KN;                            // [MAXBH][64]   newest key
  float* vn = fa + DP_A_VN;                            // [MAXBH][64]   newest value
  float* tab = fa + DP_A_TAB;                          // [MAXBH][16][DP_PART] running partial of every (pair, warp)
  float* biasrow = fa + DP_A_BIAS;                     // camera-bias row r (requested after the QKV phase; zeros without a bias)
  uint8_t* layrow = reinterpret_cast<uint8_t*>(fa + DP_A_LAY);      // [MAXBH][DP_MAXLB] layout row of query block r / lay_blk per pair
  const bool tma_bias = p.bias != nullptr && (p.bias_ld & 3) == 0;
  DP_TR(sm, 50);
  // ---- prologue: q / newest key / newest value of every pair (the latter two also appended to the cache), table reset, layout rows
  if (!tma_bias)
    for (int j = tid; j < n; j += DP_CONSUMERS) biasrow[j] = p.bias ? __ldg(p.bias + (size_t)r * p.bias_ld + j) : 0.f;
  // operands of the merge (previous residual row segment, ln1 gamma / beta of the head's channels): fetched now so that the merge does not
  // wait for L2 / HBM; the residual keeps its tag bit and is checked there.  Requested BEFORE the (polled) q / k / v loads so that both
  // are in flight together; from the last thread down, the first 384 threads fetch q / k / v.
  float* mrg = fa + DP_A_MRG;                          // [MAXBH][3][64]
  auto merge_operand = [&](int i) {
    const int k = i / 192, which = (i % 192) >> 6, c = i & 63;
    const int bh = bx + k * G, b = bh / H, h = bh - b * H;
    return which == 0 ? ldcg_f32(Xi + (size_t)b * d + h * 64 + c) : which == 1 ? __ldg(L.ln1_g + h * 64 + c) : __ldg(L.ln1_b + h * 64 + c);
  };
  const int mi0 = DP_CONSUMERS - 1 - tid;
  float mv0 = 0.f;
  if (mi0 < npairs * 192) mv0 = merge_operand(mi0);
  for (int i = tid; i < npairs * 192; i += DP_CONSUMERS) {
    const int k = i / 192, which = (i % 192) >> 6, c = i & 63;
    const int bh = bx + k * G, b = bh / H, h = bh - b * H;
    const float v = ld_tagged(sm, QKVi + (size_t)b * 3 * d + which * d + h * 64 + c, tag, 11u);
    if (which == 0) qs[k * 64 + c] = v;
    else if (which == 1) {
      kn[k * 64 + c] = v;
      reinterpret_cast<__half*>(L.kc)[(((size_t)bh * (p.Lmax >> 7) + (r >> 7)) * 64 + c) * 128 + (r & 127)] = __float2half_rn(v);
    } else {
      vn[k * 64 + c] = v;
      reinterpret_cast<__half*>(L.vc)[((size_t)bh * p.Lmax + r) * 64 + c] = __float2half_rn(v);
    }
  }
  for (int i = tid; i < npairs * 16; i += DP_CONSUMERS) { tab[i * DP_PART] = -INFINITY; tab[i * DP_PART + 1] = 0.f; }
  if (mi0 < npairs * 192) mrg[mi0] = mv0;
  for (int i = mi0 + DP_CONSUMERS; i < npairs * 192; i += DP_CONSUMERS) mrg[i] = merge_operand(i);
  if (L.layout != nullptr) {
    const int nlb = (n + p.lay_blk - 1) / p.lay_blk;
    for (int i = tid; i < npairs * nlb; i += DP_CONSUMERS) {
      const int k = i / nlb, jb = i - k * nlb;
      const int h = (bx + k * G) % H;
      layrow[k * DP_MAXLB + jb] = L.layout[((size_t)h * p.lay_ld + r / p.lay_blk) * p.lay_ld + jb];
    }
  }
  DP_TR(sm, 51);
  if (tma_bias && first_layer) {         // the row was requested at the start of the step and serves all its layers
    if (w == 0) dp_mbar_wait(sm, &sm.bias_bar, bias_par & 1u, 7u, (unsigned)n);      // one warp asks the mbarrier (see linear_phase), the barrier below tells the others
    bias_par ^= 1u;
  }
  DP_TR(sm, 52);
  bar_consumers();
  DP_TR(sm, 53);
  fine(3);
  // ---- blocks: local unit i = pair k * nblk + block -> warp group i % 4; warp q4 of the group owns keys 32 q4 .. 32 q4 + 31 of the block
  // (all 16 warps busy on four blocks at a time: the shared-memory latency is hidden by the other warps of the scheduler)
  const int grp = w >> 2, q4 = w & 3, kbase = 32 * q4;
  const int kp = lane & 15, chh = lane >> 4;           // q . K^T: lane = (key pair, channel half)
  const int kq = lane >> 3, cg = lane & 7;             // P . V:   lane = (key mod 4, channel group of 8)
  float* pwq = fa + DP_A_PW + w * 32;
  int cur = -1;
  float m_run = -INFINITY, l_run = 0.f, o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
  auto flush = [&](int k) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      o[c] += __shfl_xor_sync(0xffffffffu, o[c], 8);
      o[c] += __shfl_xor_sync(0xffffffffu, o[c], 16);
    }
    float* t = tab + (k * 16 + w) * DP_PART;
    if (lane < 8) {
      *reinterpret_cast<float4*>(t + 4 + lane * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(t + 8 + lane * 8) = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (lane == 0) { t[0] = m_run; t[1] = l_run; }
  };
  int k = 0, blk = grp;                    // unit i = pair k * nblk + block blk, advanced without a division
  while (blk >= nblk) { blk -= nblk; ++k; }
  for (int i = grp; i < nun; i += 4) {
    if (k != cur) {
      if (cur >= 0) flush(cur);
      cur = k; m_run = -INFINITY; l_run = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = 0.f;
    }
    const int j0 = blk << 7, cnt = min(128, n - j0);
    const unsigned int sq = seq + (unsigned)i;
    DP_TR(sm, 54);
    if (q4 == 0) {                         // one warp of the group waits (see linear_phase), the group barrier tells the other three
      ring_wait_prev_released(sm, sq);
      DP_TR(sm, 55);
      ring_wait_full(sm, sq);
    }
    bar_group(grp);
    DP_TR(sm, 56);
    if (kbase < cnt && !(p.dbg & 1)) {
      __half* Ks = reinterpret_cast<__half*>(sm.ring[sq % DP_NSLOT]);               // [64 channels][128 keys]
      __half* Vs = Ks + 64 * 128;                                                   // [cnt keys][64 channels]
      const int jr = r - j0;
      if (blk == nblk - 1 && jr >= kbase && jr < kbase + 32) {      // the staged copy of key r is stale: take it from the freshly computed k / v
        Ks[lane * 128 + jr] = __float2half_rn(kn[k * 64 + lane]);
        Ks[(lane + 32) * 128 + jr] = __float2half_rn(kn[k * 64 + lane + 32]);
        *reinterpret_cast<__half2*>(Vs + jr * 64 + 2 * lane) = __floats2half2_rn(vn[k * 64 + 2 * lane], vn[k * 64 + 2 * lane + 1]);
        fence_proxy_async();             // generic-proxy writes into a slot the async proxy (cp.async.bulk) refills later
        __syncwarp();
      }
      // ---- scores of keys kbase + 2 kp, + 1: each half-warp sums 32 channels
      float a0 = 0.f, a1 = 0.f;
      {
        const float4* q4p = reinterpret_cast<const float4*>(qs + k * 64 + chh * 32);
        const uint32_t* K2 = reinterpret_cast<const uint32_t*>(Ks) + (chh * 32) * 64 + 16 * q4 + kp;      // + c * 64: channel chh * 32 + c
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 qv = q4p[c4];
          const float qq[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint32_t kk = K2[(c4 * 4 + cc) * 64];
            const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(&kk));
            a0 = fmaf(qq[cc], kf.x, a0); a1 = fmaf(qq[cc], kf.y, a1);
          }
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
      }
      const int key0 = kbase + 2 * kp;
      float sc0, sc1;
      {
        const float2 b2 = *reinterpret_cast<const float2*>(biasrow + j0 + key0);
        sc0 = (a0 + b2.x) * p.scale; sc1 = (a1 + b2.y) * p.scale;
        bool ok0 = key0 < cnt, ok1 = key0 + 1 < cnt;
        if (L.layout != nullptr) {
          ok0 = ok0 && layrow[k * DP_MAXLB + (j0 + key0) / p.lay_blk] != 0;
          ok1 = ok1 && layrow[k * DP_MAXLB + (j0 + key0 + 1) / p.lay_blk] != 0;
        }
        if (!ok0) sc0 = -INFINITY;
        if (!ok1) sc1 = -INFINITY;
      }
      float mb = fmaxf(sc0, sc1);
      for (int of = 8; of; of >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, of));
      const float m_new = fmaxf(m_run, mb);
      const float p0 = (sc0 == -INFINITY) ? 0.f : expf(sc0 - m_new), p1 = (sc1 == -INFINITY) ? 0.f : expf(sc1 - m_new);
      float su = p0 + p1;
      for (int of = 8; of; of >>= 1) su += __shfl_xor_sync(0xffffffffu, su, of);
      const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
      l_run = l_run * corr + su;
      m_run = m_new;
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] *= corr;
      if (chh == 0) *reinterpret_cast<float2*>(pwq + 2 * kp) = make_float2(p0, p1);
      __syncwarp();
      // ---- P . V over the warp's keys; rows >= cnt of the slot are stale bytes and must not be touched
      {
        const uint4* V8 = reinterpret_cast<const uint4*>(Vs) + (kbase + kq) * 8 + cg;         // + ii * 32: key kbase + 4 ii + kq
        const float* pk = pwq + kq;
        auto fma8 = [&](float pv, const uint4& vv) {
          const float2 v01 = __half22float2(*reinterpret_cast<const __half2*>(&vv.x)), v23 = __half22float2(*reinterpret_cast<const __half2*>(&vv.y));
          const float2 v45 = __half22float2(*reinterpret_cast<const __half2*>(&vv.z)), v67 = __half22float2(*reinterpret_cast<const __half2*>(&vv.w));
          o[0] = fmaf(pv, v01.x, o[0]); o[1] = fmaf(pv, v01.y, o[1]); o[2] = fmaf(pv, v23.x, o[2]); o[3] = fmaf(pv, v23.y, o[3]);
          o[4] = fmaf(pv, v45.x, o[4]); o[5] = fmaf(pv, v45.y, o[5]); o[6] = fmaf(pv, v67.x, o[6]); o[7] = fmaf(pv, v67.y, o[7]);
        };
        const int nk = min(32, cnt - kbase);
        if (nk == 32) {
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) fma8(pk[4 * ii], V8[ii * 32]);
        } else {
          const int nit = (nk + 3) >> 2;
          for (int ii = 0; ii < nit; ++ii)
            if (4 * ii + kq < nk) fma8(pk[4 * ii], V8[ii * 32]);
        }
      }
    }
    __syncwarp();                         // every lane is done with the slot and with pw[]
    DP_TR(sm, 57);
    if (lane == 0) ring_release_shared(sm, sq, 4u);
    DP_TR(sm, 58);
    blk += 4;
    while (blk >= nblk) { blk -= nblk; ++k; }
  }
  if (cur >= 0) flush(cur);
  seq += (unsigned)nun;
  fine(5);
  DP_TR(sm, 59);
  bar_consumers();
  DP_TR(sm, 60);
  // The key / value appended in the prologue will be read by cp.async.bulk (async proxy) in the next step.  Here, not at the end of the
  // phase: these stores completed long ago, the ones of the merge below would make the fence wait for their round trip.
  fence_proxy_async_global();
  DP_TR(sm, 62);
  fine(9);
  // ---- merge the 16 per-warp partials of each pair (fixed order), 2 channels per lane; the residual operands are requested first
  if (w < npairs) {
    const int bh = bx + w * G, b = bh / H, h = bh - b * H;
    const int c0 = h * 64 + lane, c1 = c0 + 32;
    const size_t xi = (size_t)b * p.d;
    const float* mg = mrg + w * 192;
    float xa = mg[lane], xb = mg[lane + 32];               // staged by the prologue; their tags are checked after the merge arithmetic
    const float ga = mg[64 + lane], gb = mg[96 + lane], ba = mg[128 + lane], bb = mg[160 + lane];
    const float* t = tab + (w * 16) * DP_PART;
    // lane i < 16 owns partial i: its weight exp(m_i - M) is computed once and broadcast
    const float mi = (lane < 16) ? t[lane * DP_PART] : -INFINITY;
    float M = mi;
    for (int of = 8; of; of >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, of));
    M = __shfl_sync(0xffffffffu, M, 0);
    const float wi = (mi == -INFINITY) ? 0.f : expf(mi - M);
    const float li = (lane < 16) ? t[lane * DP_PART + 1] * wi : 0.f;
    float Ls = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int ww = 0; ww < 16; ++ww) {
      const float wgt = __shfl_sync(0xffffffffu, wi, ww);
      Ls += __shfl_sync(0xffffffffu, li, ww);
      if (wgt != 0.f) {                   // a warp without a block of this pair left whatever the scratch held in o[]
        o0 += t[ww * DP_PART + 4 + lane] * wgt;
        o1 += t[ww * DP_PART + 4 + 32 + lane] * wgt;
      }
    }
    if (((__float_as_uint(xa) ^ ptag) | (__float_as_uint(xb) ^ ptag)) & 1u) {
      DpPoll pg;
      do { poll_tick(sm, pg, 12u, (unsigned)bh); xa = ldcg_f32(Xi + xi + c0); xb = ldcg_f32(Xi + xi + c1); } while (((__float_as_uint(xa) ^ ptag) | (__float_as_uint(xb) ^ ptag)) & 1u);
    }
    __syncwarp();
    const float2 st = sm.rowstat[0][b];
    const float x0 = ((xa - st.x) * st.y * ga + ba) + o0 / Ls, x1 = ((xb - st.x) * st.y * gb + bb) + o1 / Ls;
    X1o[xi + c0] = tag_f32(x0, tag);
    X1o[xi + c1] = tag_f32(x1, tag);
    store_frag(X1Fo, b, c0, x0, tag);
    store_frag(X1Fo, b, c1, x1, tag);
    float sv = x0 + x1, qv = x0 * x0 + x1 * x1;
    for (int of = 16; of; of >>= 1) { sv += __shfl_xor_sync(0xffffffffu, sv, of); qv += __shfl_xor_sync(0xffffffffu, qv, of); }
    if (lane == 0) *reinterpret_cast<float2*>(PSX1o + ((size_t)b * p.H + h) * 2) = make_float2(tag_f32(sv, tag), tag_f32(qv, tag));
  }
  DP_TR(sm, 61);
  fine(6);
}

// ------------------------------------------------------------------------------------------------ sampling + embedding (CTA b < B)
__device__ __forceinline__ void dp_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
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

// Same tail as dec_sample_kernel (decode.cu): logits / T, top-k keeping ties with the k-th value, softmax, Philox multinomial | greedy |
// forced token (cond_transformer_multi_view.py:138-142,200-219).  Returns the token (uniform over the consumer threads).
__device__ __noinline__ int sample_row(const DecodeParams& p, DpSmem& sm, int b, int s) {
  const int tid = threadIdx.x, V = p.vocab;
  float* lg = reinterpret_cast<float*>(sm.act);
  int* hist = reinterpret_cast<int*>(sm.act) + DP_MAXV;          // 256 bins of the radix select
  const float inv_t = 1.0f / p.temperature;
  for (int i = tid; i < V; i += DP_CONSUMERS) {
    float v = __ldcg(p.LOGITS + (size_t)b * V + i);
    if (p.trace != nullptr) p.trace[((size_t)s * p.B + b) * V + i] = v;
    lg[i] = v * inv_t;
  }
  bar_consumers();
  float thr = -INFINITY;
  if (p.top_k > 0 && p.top_k < V) {
    // k-th largest value by a 4-pass, 8-bit radix select on the order-preserving integer image of the floats (exactly the value a
    // descending sort would leave at position k - 1, so ties with it are kept like the reference's `out < v[..., [-1]]` filter)
    unsigned int prefix = 0u, known = 0u;
    int krem = p.top_k;
    for (int pass = 3; pass >= 0; --pass) {
      for (int i = tid; i < 256; i += DP_CONSUMERS) hist[i] = 0;
      bar_consumers();
      for (int i = tid; i < V; i += DP_CONSUMERS) {
        const unsigned int u = __float_as_uint(lg[i]);
        const unsigned int key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        if ((key & known) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1);
      }
      bar_consumers();
      if (tid < 32) {                      // lane owns bins 8 lane .. 8 lane + 7; suffix sums run from the top bin down
        int c[8], local = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) { c[e] = hist[tid * 8 + e]; local += c[e]; }
        int incl = local;                  // elements in this lane's bins and all higher lanes' bins
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_down_sync(0xffffffffu, incl, o);
          if (tid + o < 32) incl += t;
        }
        const int above = incl - local;
        if (above < krem && krem <= incl) {
          int acc = above;
#pragma unroll
          for (int e = 7; e >= 0; --e) {
            if (acc + c[e] >= krem) { sm.flags[0] = (unsigned)(tid * 8 + e); sm.flags[1] = (unsigned)(krem - acc); break; }
            acc += c[e];
          }
        }
      }
      bar_consumers();
      prefix |= sm.flags[0] << (8 * pass);
      known |= 255u << (8 * pass);
      krem = (int)sm.flags[1];
    }
    thr = __uint_as_float((prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix);
  }
  float m = -INFINITY;
  for (int i = tid; i < V; i += DP_CONSUMERS) m = fmaxf(m, lg[i]);
  m = consumers_max(m, sm.red16);
  const int per = (V + DP_CONSUMERS - 1) / DP_CONSUMERS;
  float local = 0.f;
  for (int e = 0; e < per; ++e) {
    const int i = tid * per + e;
    if (i < V) {
      const float pr = (lg[i] >= thr) ? expf(lg[i] - m) : 0.f;
      lg[i] = pr;
      local += pr;
    }
  }
  const float total = consumers_sum(local, sm.red16);
  const long long fz = (p.forced != nullptr) ? p.forced[(size_t)b * p.n_img + s] : -1;
  int token = 0;
  if (fz >= 0) {
    token = (int)fz;
  } else if (p.greedy) {
    float best = -1.f;
    int bi = 0;
    for (int i = tid; i < V; i += DP_CONSUMERS)
      if (lg[i] > best) { best = lg[i]; bi = i; }
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    bar_consumers();
    if ((tid & 31) == 0) { sm.red16[tid >> 5] = best; sm.flags[tid >> 5] = (unsigned)bi; }
    bar_consumers();
    best = sm.red16[0]; bi = (int)sm.flags[0];
    for (int ww = 1; ww < 16; ++ww)
      if (sm.red16[ww] > best || (sm.red16[ww] == best && (int)sm.flags[ww] < bi)) { best = sm.red16[ww]; bi = (int)sm.flags[ww]; }
    token = bi;
  } else {
    uint32_t rnd[4];
    dp_philox((uint32_t)s, (uint32_t)b, 0u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), rnd);
    const float u = ((rnd[0] >> 8) + 0.5f) * (1.0f / 16777216.0f) * total;
    float incl = local;
    for (int o = 1; o < 32; o <<= 1) {
      const float tt = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += tt;
    }
    bar_consumers();
    if ((tid & 31) == 31) sm.red16[tid >> 5] = incl;
    if (tid == 0) sm.found = V;
    bar_consumers();
    float base = 0.f;
    for (int ww = 0; ww < (tid >> 5); ++ww) base += sm.red16[ww];
    float run = base + incl - local;
    for (int e = 0; e < per; ++e) {
      const int i = tid * per + e;
      if (i < V && lg[i] > 0.f) {
        run += lg[i];
        if (run > u) { atomicMin(&sm.found, i); break; }
      }
    }
    bar_consumers();
    token = sm.found;
    if (token >= V) {
      int last = 0;
      for (int i = 0; i < V; ++i) if (lg[i] > 0.f) last = i;
      token = last;
    }
  }
  bar_consumers();
  return token;
}

// Embedding of decode-order image token `sdec` (value tok) of scene b (mingpt_sparse.py:332-350; same arithmetic as embed_kernel)
// -> the residual-stream row in fp32 (X), in fragment order (XF) and its LayerNorm sums (PSX: the whole row in part 0, zeros elsewhere).
// Written as generation `c` (the layer before the first layer of the step that consumes it; real grid barriers surround this phase).
__device__ __noinline__ void embed_row(const DecodeParams& p, DpSmem& sm, int b, int sdec, long long tok, int c) {
  const int tid = threadIdx.x, d = p.d;
  const unsigned int tag = dp_tag(c);
  uint8_t* XFo = p.XF + (size_t)dp_inst(c) * (size_t)(d >> 6) * 4096;
  const int j = p.fwd[sdec];
  const int cam = j / p.hw, px = j - cam * p.hw;
  const float* e = p.x_tok_emb + (size_t)tok * d;
  const float* pos = p.x_pos_emb + (size_t)j * d;
  float* out = p.X + (size_t)dp_inst(c) * 16 * d + (size_t)b * d;
  float gv[2] = {0.f, 0.f};
  float inv = 0.f;
  if (p.img_embed_w != nullptr) {
    const float* I = p.I_inv + ((size_t)b * p.ncam + cam) * 9;
    const float* E = p.E_inv + ((size_t)b * p.ncam + cam) * 16;
    const float* pix = p.pixel + (size_t)px * 3;
    float cv[4], ray[4];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) cv[rr] = I[rr * 3 + 0] * pix[0] + I[rr * 3 + 1] * pix[1] + I[rr * 3 + 2] * pix[2];
    cv[3] = 1.0f;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) ray[rr] = E[rr * 4 + 0] * cv[0] + E[rr * 4 + 1] * cv[1] + E[rr * 4 + 2] * cv[2] + E[rr * 4 + 3] * cv[3];
    float ss = 0.f;
    int k = 0;
    for (int ch = tid; ch < d; ch += DP_CONSUMERS, ++k) {
      const float4 wi = __ldg(reinterpret_cast<const float4*>(p.img_embed_w) + ch);
      const float4 wc = __ldg(reinterpret_cast<const float4*>(p.cam_embed_w) + ch);
      const float de = wi.x * ray[0] + wi.y * ray[1] + wi.z * ray[2] + wi.w * ray[3];
      const float ce = wc.x * E[3] + wc.y * E[7] + wc.z * E[11] + wc.w * E[15];
      gv[k] = de - ce;
      ss += gv[k] * gv[k];
    }
    ss = consumers_sum(ss, sm.red16);
    inv = 1.0f / (sqrtf(ss) + 1e-7f);
  }
  float sv = 0.f, qv = 0.f;
  int k = 0;
  for (int ch = tid; ch < d; ch += DP_CONSUMERS, ++k) {
    const float x = (e[ch] + gv[k] * inv) + pos[ch];
    out[ch] = tag_f32(x, tag);
    store_frag(XFo, b, ch, x, tag);
    sv += x; qv += x * x;
  }
  sv = consumers_sum(sv, sm.red16);
  qv = consumers_sum(qv, sm.red16);
  const int nparts = d >> 3;
  float2* ps = reinterpret_cast<float2*>(p.PSX + (size_t)dp_inst(c) * 16 * nparts * 2) + (size_t)b * nparts;
  for (int i = tid; i < nparts; i += DP_CONSUMERS) ps[i] = (i == 0) ? make_float2(tag_f32(sv, tag), tag_f32(qv, tag)) : make_float2(tag_f32(0.f, tag), tag_f32(0.f, tag));
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(DP_THREADS, 1) decode_persistent_kernel(const __grid_constant__ DecodeParams p) {
  extern __shared__ __align__(1024) uint8_t dp_raw[];
  DpSmem& sm = *reinterpret_cast<DpSmem*>(dp_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < DP_NSLOT; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
    mbar_init(&sm.act_bar, 1);
    mbar_init(&sm.bias_bar, 1);
    fence_barrier_init();
    sm.debug = p.debug;
    part_range(3 * p.d / 8, sm.rng[RNG_QKV][0], sm.rng[RNG_QKV][1]);
    part_range(4 * p.d / 8, sm.rng[RNG_MLP1][0], sm.rng[RNG_MLP1][1]);
    quarter_range(p.d / 8, sm.rng[RNG_MLP2][0], sm.rng[RNG_MLP2][1]);
    part_range(p.vpad / 8, sm.rng[RNG_HEAD][0], sm.rng[RNG_HEAD][1]);
    sm.trace = (p.profile != nullptr && (p.dbg & 64) && blockIdx.x == (unsigned)p.trace_cta) ? p.profile + (size_t)gridDim.x * 32 : nullptr;
    sm.trace_n = 0; sm.trace_on = 0;
    sm.where[0] = sm.where[1] = sm.where[2] = 0u;
    for (int i = 0; i < DP_NSLOT; ++i) { sm.rel[i] = 0xffffffffu; sm.slotcnt[i] = 0u; }
    for (int i = 0; i < 20; ++i) sm.fine[i] = 0ull;
    for (int i = 0; i < 12; ++i) sm.prof[i] = 0ull;
    sm.att_epoch = 0u;
  }
  __syncthreads();
  if (tid >= DP_CONSUMERS) {
    if (tid == DP_CONSUMERS) dp_producer(p, sm);
    return;
  }
  unsigned int seq = 0, bar_target = 0, act_par = 0, bias_par = 0;
  const unsigned int G = gridDim.x;
  const int d = p.d, H = p.H;
  const bool has_pairs = (int)blockIdx.x < p.B * H;
  const bool tma_bias = p.bias != nullptr && (p.bias_ld & 3) == 0;
  // X <- embedding of the token drawn at step step_begin - 1 (it is already in the token grid), as generation -1
  if ((int)blockIdx.x < p.B) {
    const int b = blockIdx.x, sdec = p.step_begin - 1;
    const int j = p.fwd[sdec];
    embed_row(p, sm, b, sdec, p.cam_idx[((size_t)b * p.ncam + j / p.hw) * p.hw + (j % p.hw)], -1);
  }
  grid_sync(sm, p.barrier, bar_target, G);
  // optional per-CTA profile: nanoseconds spent in each phase body and in each phase boundary, summed over the launch
  unsigned long long tmark = dp_globaltimer();
  auto mark = [&](int slot, unsigned int step, unsigned int layer, unsigned int phase) {
    if (tid == 0) {
      if (p.profile != nullptr) {
        const unsigned long long now = dp_globaltimer();
        sm.prof[slot] += now - tmark;
        tmark = now;
      }
      sm.where[0] = step; sm.where[1] = layer; sm.where[2] = phase;
    }
  };
  const size_t vecb = (size_t)(d >> 6) * 4096;
  int c = 0;                               // running layer number of this launch: instance c & 1, tag dp_tag(c) of everything layer c writes
  for (int s = p.step_begin; s < p.step_end; ++s) {
    if (tid == 0 && has_pairs && tma_bias) {
      // camera-bias row of this step (the same for every layer) -> upper half of act[]: not part of any reduction / sampling scratch and only
      // ever written by these copies; its readers of the previous step are behind the grid barriers.  Waited for in the first attention phase.
      const int n = p.nc + s;
      const uint32_t bytes = (uint32_t)((n + 3) & ~3) * 4u;
      mbar_expect_tx(&sm.bias_bar, bytes);
      dp_bulk_g2s(sm.act + DP_A_BIAS * 4, p.bias + (size_t)(n - 1) * p.bias_ld, bytes, &sm.bias_bar);
    }
    for (int l = 0; l < p.n_layers; ++l, ++c) {
      const DecodeLayer& L = p.layers[l];
      // (a copy of the layer block in shared memory measured 3 % slower; the global block only gets an L2 prefetch one layer ahead, it falls
      // out of L2 between two token steps)
      if (tid == DP_CONSUMERS - 32) {
        const char* nl = reinterpret_cast<const char*>(p.layers + (l + 1 == p.n_layers ? 0 : l + 1));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nl));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nl + sizeof(DecodeLayer) - 4));
      }
      if (tid == 0 && sm.trace != nullptr) sm.trace_on = (s == p.trace_step && l == p.n_layers / 4) ? 1 : 0;
      DP_TR(sm, 1);
      mark(11, s, l, 0);
      linear_phase(p, sm, seq, act_par, EPI_QKV, p.XF + (size_t)dp_inst(c - 1) * vecb, p.PSX + (size_t)dp_inst(c - 1) * 16 * (d >> 3) * 2, d >> 3, 0, RNG_QKV,
                   L.s_qkv, L.c1_qkv, L.c2_qkv, dp_tag(c - 1), p.QKV + (size_t)dp_inst(c) * 16 * 3 * d, dp_tag(c));
      mark(0, s, l, 1);
      DP_TR(sm, 2);
      phase_sync(sm);
      mark(1, s, l, 2);
      DP_TR(sm, 3);
      attention_phase(p, sm, seq, bias_par, L, s, c, l == 0);
      mark(2, s, l, 3);
      DP_TR(sm, 4);
      phase_sync(sm);
      // this CTA's appends of (s, l) are done (it owns its pairs' cache rows): the producer may request the blocks of (s + 1, l)
      // (published by a thread that has no global stores of the merge in flight: the fence does not wait)
      if (tid == DP_CONSUMERS - 1) { __threadfence_block(); sm.att_epoch = (unsigned)(s - p.step_begin) * (unsigned)p.n_layers + (unsigned)l + 1u; }
      mark(3, s, l, 4);
      DP_TR(sm, 5);
      linear_phase(p, sm, seq, act_par, EPI_MLP1, p.X1F + (size_t)dp_inst(c) * vecb, p.PSX1 + (size_t)dp_inst(c) * 16 * H * 2, H, 1, RNG_MLP1, L.s_1, L.c1_1,
                   L.c2_1, dp_tag(c), p.HF + (size_t)dp_inst(c) * 4 * vecb, dp_tag(c));
      mark(4, s, l, 5);
      DP_TR(sm, 6);
      phase_sync(sm);
      mark(5, s, l, 6);
      DP_TR(sm, 7);
      {
        const int cpq = (int)G >> 2, q2 = min((int)blockIdx.x / cpq, 3);
        linear_phase(p, sm, seq, act_par, EPI_M2P, p.HF + ((size_t)dp_inst(c) * 4 + q2) * vecb, nullptr, 0, 0, RNG_MLP2, L.s_2, nullptr, nullptr, dp_tag(c),
                     p.P2 + (size_t)dp_inst(c) * 4 * 16 * d, dp_tag(c));
        mlp2_finalize(p, sm, L, c);
      }
      mark(6, s, l, 7);
      DP_TR(sm, 8);
      phase_sync(sm);
      mark(7, s, l, 8);
      DP_TR(sm, 9);
      if (tid == 0 && sm.trace_on) { sm.trace_on = 0; sm.trace[1023] = (unsigned long long)sm.trace_n; }
    }
    linear_phase(p, sm, seq, act_par, EPI_HEAD, p.XF + (size_t)dp_inst(c - 1) * vecb, p.PSX + (size_t)dp_inst(c - 1) * 16 * (d >> 3) * 2, d >> 3, 0, RNG_HEAD,
                 p.s_head, p.c1_head, p.c2_head, dp_tag(c - 1), nullptr, 0u);
    mark(8, s, p.n_layers, 9);
    grid_sync(sm, p.barrier, bar_target, G);
    mark(9, s, p.n_layers, 10);
    if ((int)blockIdx.x < p.B) {
      const int b = blockIdx.x;
      const int token = sample_row(p, sm, b, s);
      if (tid == 0) {
        const int j = p.fwd[s];
        p.cam_idx[((size_t)b * p.ncam + j / p.hw) * p.hw + (j % p.hw)] = token;
        if (p.tokens_out != nullptr) p.tokens_out[(size_t)b * p.n_img + s] = token;
      }
      if (s + 1 < p.step_end) embed_row(p, sm, b, s, token, c - 1);
    }
    mark(10, s, p.n_layers, 11);
    grid_sync(sm, p.barrier, bar_target, G);
  }
  if (tid == 0 && p.profile != nullptr) {
#pragma unroll
    for (int i = 0; i < 12; ++i) p.profile[(size_t)blockIdx.x * 32 + i] = sm.prof[i];
    for (int i = 0; i < 10; ++i) p.profile[(size_t)blockIdx.x * 32 + 12 + i] = sm.fine[i];
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// W [n_rows][ld] fp32 (row-major, K contiguous) -> units of 8 rows x d columns in mma B-fragment order (see the header of this file):
// unit u = quarter * ceil8(n_rows) / 8 + row_unit covers columns quarter * d .. quarter * d + d - 1.  One thread per 16-byte chunk.
__global__ void pack_decode_linear_kernel(const float* __restrict__ W, int n_rows, int ld, int d, int n_quarters, float lo_mul,
                                          uint8_t* __restrict__ out, long long n_chunks) {
  const long long ci = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= n_chunks) return;
  const int KG = d >> 6;
  const int lane = (int)(ci & 31);
  const int part = (int)((ci >> 5) % 3);
  const long long ukg = (ci >> 5) / 3;
  const int kg = (int)(ukg % KG);
  const long long u = ukg / KG;
  const int units_per_q = (n_rows + 7) >> 3;
  const int q = (int)(u / units_per_q), ru = (int)(u % units_per_q);
  const int g = lane >> 2, t = lane & 3;
  const int row = ru * 8 + g;
  const float* wr = W + (size_t)row * ld + (size_t)q * d + kg * 64;
  const bool ok = row < n_rows;
  uint4 o;
  if (part < 2) {
    uint32_t r[4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k0 = (2 * part + e) * 16 + 2 * t;
      const float a = ok ? wr[k0] : 0.f, b = ok ? wr[k0 + 1] : 0.f, c = ok ? wr[k0 + 8] : 0.f, dd = ok ? wr[k0 + 9] : 0.f;
      r[2 * e] = (uint32_t)__half_as_ushort(__float2half_rn(a)) | ((uint32_t)__half_as_ushort(__float2half_rn(b)) << 16);
      r[2 * e + 1] = (uint32_t)__half_as_ushort(__float2half_rn(c)) | ((uint32_t)__half_as_ushort(__float2half_rn(dd)) << 16);
    }
    o = make_uint4(r[0], r[1], r[2], r[3]);
  } else {
    uint32_t r[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int k0 = ks * 16 + 2 * t;
      const int kk[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
      uint32_t v = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float wv = ok ? wr[kk[e]] : 0.f;
        const float res = (wv - __half2float(__float2half_rn(wv))) * lo_mul;
        v |= (uint32_t)__nv_cvt_float_to_fp8(res, __NV_SATFINITE, __NV_E4M3) << (8 * e);
      }
      r[ks] = v;
    }
    o = make_uint4(r[0], r[1], r[2], r[3]);
  }
  reinterpret_cast<uint4*>(out)[ci] = o;
}

int launch_pack_decode_linear(const float* W, int n_rows, int ld, int d, int n_quarters, float lo_mul, void* out, cudaStream_t st) {
  if (d % 64 != 0 || d < 64 || d > 1024 || n_rows < 1 || n_quarters < 1 || ld < n_quarters * d) return BEVGEN_ERR_ARG;
  const long long units = (long long)n_quarters * ((n_rows + 7) / 8);
  const long long chunks = units * (d / 64) * 96;
  pack_decode_linear_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(W, n_rows, ld, d, n_quarters, lo_mul, (uint8_t*)out, chunks);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

long long decode_packed_bytes(int n_rows, int d, int n_quarters) {
  return (long long)n_quarters * ((n_rows + 7) / 8) * (d / 64) * DP_KG_BYTES;
}

void decode_workspace_sizes(int B, int d, int H, int vocab, long long* n_floats, long long* n_counters) {
  const long long vpad = (vocab + 7) / 8 * 8;
  // every buffer that crosses CTAs inside a step exists twice (instance = running layer number & 1, see "tag sync")
  *n_floats = 2 * (16LL * d * 2 /*X, X1*/ + 16LL * 3 * d /*QKV*/) + 16LL * vpad /*LOGITS*/ + 2 * (16LL * (d / 8) * 2 + 16LL * H * 2) /*PSX, PSX1*/ +
              2 * (16LL * d * 2 /*XF, X1F*/ + 4LL * 16 * d /*HF*/) + 2 * 4LL * 16 * d /*P2*/;
  *n_counters = 64;
  (void)B;
}

int launch_decode_persistent(DecodeParams p, float* ws, unsigned int* counters, int sm_count, cudaStream_t st) {
  const int d = p.d;
  if (p.B < 1 || p.B > 16 || d % 64 != 0 || d < 64 || d > 1024 || p.H * 64 != d || p.vocab < 1 || p.vocab > DP_MAXV || p.Lmax > DP_MAXL || (p.Lmax & 127) ||
      p.n_layers < 1 || p.step_begin < 1 || p.step_end > p.n_img || p.step_begin > p.step_end || p.nc + p.n_img > p.Lmax || p.temperature <= 0.f)
    return BEVGEN_ERR_ARG;
  if (p.step_begin == p.step_end) return BEVGEN_OK;
  const int G = sm_count;
  if ((p.B * p.H + G - 1) / G > DP_MAXBH) return BEVGEN_ERR_ARG;
  if (G < 4 || G < d / 8) return BEVGEN_ERR_ARG;      // MLP2: four K-quarter groups of CTAs, and CTA ru finishes row unit ru < d / 8
  for (int l = 0; l < 1; ++l)
    if (p.lay_ld > 0 && (p.lay_blk < 16 || (p.Lmax + p.lay_blk - 1) / p.lay_blk > DP_MAXLB)) return BEVGEN_ERR_ARG;
  p.vpad = (p.vocab + 7) / 8 * 8;
  if (const char* e = getenv("BEVGEN_DP_DBG")) p.dbg = atoi(e);
  p.trace_cta = 0; p.trace_step = p.step_begin + 300;
  if (const char* e = getenv("BEVGEN_DP_TRACE_CTA")) p.trace_cta = atoi(e);
  if (const char* e = getenv("BEVGEN_DP_TRACE_STEP")) p.trace_step = atoi(e);      // timing experiments only (results are garbage): see tools/decode_debug.py
  float* f = ws;
  p.X = f; f += 2 * 16 * d;                                   // [2 instances][16][d]
  p.X1 = f; f += 2 * 16 * d;
  p.QKV = f; f += 2 * 16 * 3 * d;
  p.LOGITS = f; f += 16 * p.vpad;
  p.PSX = f; f += 2 * 16 * (d / 8) * 2;
  p.PSX1 = f; f += 2 * 16 * p.H * 2;
  p.XF = reinterpret_cast<uint8_t*>(f); f += 2 * 16 * d;      // [2] x 16 x d x (2 + 2) bytes
  p.X1F = reinterpret_cast<uint8_t*>(f); f += 2 * 16 * d;
  p.HF = reinterpret_cast<uint8_t*>(f); f += 2 * 4 * 16 * d;
  p.P2 = f; f += 2 * 4 * 16 * d;                              // [2][4 K-quarters][16][d] partial sums of MLP2
  p.barrier = counters;
  // all tags start at 0: the first generation written to every instance carries tag 1 (dp_tag), leftovers of an earlier launch must not pass
  if (cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)(f - ws), st) != cudaSuccess) return BEVGEN_ERR_CUDA;
  if (cudaMemsetAsync(counters, 0, sizeof(unsigned int) * 64, st) != cudaSuccess) return BEVGEN_ERR_CUDA;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DpSmem)) != cudaSuccess) return BEVGEN_ERR_CUDA;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(DP_THREADS);
  cfg.dynamicSmemBytes = sizeof(DpSmem);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident (the grid barrier depends on it)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, decode_persistent_kernel, p) != cudaSuccess) return BEVGEN_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
