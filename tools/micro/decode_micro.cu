// Micro-benchmarks that calibrate the persistent decode kernel's design (run on the GPU box: tools/micro/decode_micro):
//  1. legacy mma.sync m16n8k16 f16 -> f32 issue rate per SM (16 warps, 4 independent accumulator chains each)
//  2. grid barrier latency over 148 co-resident CTAs: acquire-poll vs relaxed-poll + fence
//  3. 148 CTAs reading the same 64 KB from L2 with ld.global.cg (the activation broadcast of every phase)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void k_hmma(float* out, long long* cyc, int iters) {
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
  uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    mma_f16(c0, a, b0, b1); mma_f16(c1, a, b0, b1); mma_f16(c2, a, b0, b1); mma_f16(c3, a, b0, b1);
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[2] + c3[3];
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned* p) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

template <int MODE>
__global__ void k_barrier(unsigned* ctr, long long* cyc, int rounds, float* sink) {
  unsigned target = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      if (MODE == 0) { red_release(ctr); while (ld_acquire(ctr) < target) {} }
      else if (MODE == 1) { red_release(ctr); while (ld_relaxed(ctr) < target) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
      else { __threadfence(); red_relaxed(ctr); while (ld_relaxed(ctr) < target) {} __threadfence(); }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_bcast(const float* src, unsigned* ctr, long long* cyc, int rounds, float* sink, int floats_per_cta) {
  unsigned target = 0;
  float acc = 0.f;
  long long total = 0;
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) { red_release(ctr); while (ld_acquire(ctr) < target) {} }
    __syncthreads();
    long long t0 = clock64();
    for (int i = threadIdx.x * 2; i < floats_per_cta; i += blockDim.x * 2) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(src + i));
      acc += v.x + v.y;
    }
    __syncthreads();
    total += clock64() - t0;
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = total;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// variants of the broadcast read: MODE 0 = same order (float2), 1 = start rotated by the CTA index (float2), 2 = rotated float4,
// 3 = one cp.async.bulk of the whole vector into shared memory (TMA engine), 4 = bulk copy in 16 rotated 1/16th pieces
template <int MODE>
__global__ void k_bcast2(const float* src, unsigned* ctr, long long* cyc, int rounds, float* sink, int floats_per_cta) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned target = 0;
  float acc = 0.f;
  long long total = 0;
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), sm_a = (unsigned)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) { red_release(ctr); while (ld_acquire(ctr) < target) {} }
    __syncthreads();
    long long t0 = clock64();
    if (MODE <= 1) {
      const int rot = MODE == 1 ? (int)((blockIdx.x * 2654435761u) % (unsigned)floats_per_cta) & ~63 : 0;
      for (int i = threadIdx.x * 2; i < floats_per_cta; i += blockDim.x * 2) {
        int j = i + rot; if (j >= floats_per_cta) j -= floats_per_cta;
        const float2 v = __ldcg(reinterpret_cast<const float2*>(src + j));
        acc += v.x + v.y;
      }
    } else if (MODE == 2) {
      const int rot = (int)((blockIdx.x * 2654435761u) % (unsigned)floats_per_cta) & ~63;
      for (int i = threadIdx.x * 4; i < floats_per_cta; i += blockDim.x * 4) {
        int j = i + rot; if (j >= floats_per_cta) j -= floats_per_cta;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + j));
        acc += v.x + v.y + v.z + v.w;
      }
    } else {
      const unsigned bytes = (unsigned)floats_per_cta * 4u;
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        if (MODE == 3) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_a), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
        } else {
          const unsigned piece = bytes / 16u;
          for (unsigned k = 0; k < 16u; ++k) {
            const unsigned kk = (k + blockIdx.x) & 15u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_a + kk * piece),
                         "l"(reinterpret_cast<const unsigned char*>(src) + kk * piece), "r"(piece), "r"(bar_a) : "memory");
          }
        }
      }
      unsigned ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a), "r"((unsigned)(r & 1)) : "memory");
      acc += reinterpret_cast<float*>(smem)[threadIdx.x];
    }
    __syncthreads();
    total += clock64() - t0;
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = total;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; long long* cyc; unsigned* ctr; float* src;
  cudaMalloc(&out, sizeof(float) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms); cudaMalloc(&ctr, 4); cudaMalloc(&src, 1 << 20);
  cudaMemset(src, 0, 1 << 20);
  std::vector<long long> h(sms);
  auto avg = [&]() { cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost); double s = 0; for (auto v : h) s += v; return s / sms; };
  for (int warps : {4, 8, 16}) {
    const int iters = 4096;
    k_hmma<<<sms, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k_hmma<<<sms, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    const double c = avg();
    printf("hmma m16n8k16 f16->f32: %2d warps/SM: %.2f cycles per MMA per SM (%.1f per warp-MMA)\n", warps, c / (iters * 4.0 * warps), c / (iters * 4.0));
  }
  const int rounds = 2000;
  void* args0[] = {&ctr, &cyc, (void*)&rounds, &out};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(ctr, 0, 4);
      const void* fn = mode == 0 ? (const void*)k_barrier<0> : mode == 1 ? (const void*)k_barrier<1> : (const void*)k_barrier<2>;
      cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(512), args0, 0, 0);
      cudaDeviceSynchronize();
    }
    printf("grid barrier mode %d (%s): %.0f cycles per barrier (512 threads/CTA, %d CTAs)  err=%s\n", mode,
           mode == 0 ? "red.release + ld.acquire poll" : mode == 1 ? "red.release + relaxed poll + fence" : "threadfence + relaxed", avg() / rounds, sms,
           cudaGetErrorString(cudaGetLastError()));
  }
  for (int kb : {16, 64}) {
    int fl = kb * 256;
    void* args1[] = {&src, &ctr, &cyc, (void*)&rounds, &out, &fl};
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(ctr, 0, 4);
      cudaLaunchCooperativeKernel((const void*)k_bcast, dim3(sms), dim3(512), args1, 0, 0);
      cudaDeviceSynchronize();
    }
    printf("L2 broadcast read of the same %d KB by all %d CTAs (ld.cg float2, 512 threads): %.0f cycles per round  err=%s\n", kb, sms, avg() / rounds,
           cudaGetErrorString(cudaGetLastError()));
  }
  {
    int fl = 64 * 256;
    void* args2[] = {&src, &ctr, &cyc, (void*)&rounds, &out, &fl};
    const void* fns[5] = {(const void*)k_bcast2<0>, (const void*)k_bcast2<1>, (const void*)k_bcast2<2>, (const void*)k_bcast2<3>, (const void*)k_bcast2<4>};
    const char* names[5] = {"same order float2", "rotated start float2", "rotated start float4", "one cp.async.bulk (TMA)", "16 rotated cp.async.bulk pieces"};
    for (int m = 0; m < 5; ++m) {
      cudaFuncSetAttribute(fns[m], cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(ctr, 0, 4);
        cudaLaunchCooperativeKernel(fns[m], dim3(sms), dim3(512), args2, 65536, 0);
        cudaDeviceSynchronize();
      }
      printf("L2 broadcast 64 KB, %-32s: %.0f cycles per round  err=%s\n", names[m], avg() / rounds, cudaGetErrorString(cudaGetLastError()));
    }
  }
  printf("clock %d kHz, %d SMs\n", clk, sms);
  return 0;
}
