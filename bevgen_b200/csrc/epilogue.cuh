// Coalesced epilogue helper shared by the tcgen05 kernels.
// After tcgen05.ld every thread owns ONE output row (pixel) and 32 consecutive columns: storing that directly makes each warp
// instruction touch 32 different 128-byte lines.  Here the 32x32 fp32 chunk is transposed through a per-warp 4 KB shared-memory
// buffer (16-byte units XOR-swizzled by row, conflict-free both ways) so that every global load/store instruction of the warp moves
// four complete 128-byte lines: lane l handles row 4*i + (l >> 3), columns 4*(l & 7) .. +3 for i = 0..7.
// The residual add and the GroupNorm statistics of the final value happen in that transposed domain.
#pragma once
#include "common.cuh"

namespace bevgen {

struct RowMap {            // global element offsets of the 8 rows this lane touches (4*i + (lane >> 3)), -1 = outside the image
  long long off[8];
};

// cpg = channels per GroupNorm group (2, 4, 8, 16 or 32); wacc = this warp's fp64 accumulators [64] (group-major: sum, sumsq)
__device__ __forceinline__ void epilogue_chunk32(const float (&v)[32], uint8_t* stage, int lane, const RowMap& rm, int col0 /*within row*/,
                                                 const float* __restrict__ residual, float* __restrict__ out, int cpg, int gcol0,
                                                 double* wacc) {
  // stage: thread = row
  {
    uint8_t* prow = stage + lane * 128;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      *reinterpret_cast<float4*>(prow + ((u ^ (lane & 7)) * 16)) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
  }
  __syncwarp();
  const int unit = lane & 7, rsub = lane >> 3;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;          // (sum, sumsq) of channels {0,1} and {2,3} of this lane's unit
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = 4 * i + rsub;
    float4 x = *reinterpret_cast<const float4*>(stage + row * 128 + ((unit ^ (row & 7)) * 16));
    const long long o = rm.off[i];
    if (o >= 0) {
      if (residual != nullptr) {
        const float4 r4 = *reinterpret_cast<const float4*>(residual + o + col0 + unit * 4);
        x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
      }
      *reinterpret_cast<float4*>(out + o + col0 + unit * 4) = x;
      s0 += x.x + x.y; q0 += x.x * x.x + x.y * x.y;
      s1 += x.z + x.w; q1 += x.z * x.z + x.w * x.w;
    }
  }
  __syncwarp();
  if (wacc == nullptr) return;
  if (cpg >= 4) { s0 += s1; q0 += q1; }
  // rows: lanes with the same unit (xor 8, 16)
  s0 += __shfl_xor_sync(0xffffffffu, s0, 8);  q0 += __shfl_xor_sync(0xffffffffu, q0, 8);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
  if (cpg == 2) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, 8);  q1 += __shfl_xor_sync(0xffffffffu, q1, 8);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
    if (lane < 8) {
      const int g = gcol0 / 2 + lane * 2;
      wacc[g * 2] += (double)s0; wacc[g * 2 + 1] += (double)q0;
      wacc[g * 2 + 2] += (double)s1; wacc[g * 2 + 3] += (double)q1;
    }
    return;
  }
  // wider groups: merge neighbouring units
  for (int w = 1; w < cpg / 4; w <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, w);
    q0 += __shfl_xor_sync(0xffffffffu, q0, w);
  }
  const int upg = cpg / 4;                                 // units per group
  if (lane < 8 && (lane % upg) == 0) {
    const int g = gcol0 / cpg + lane / upg;
    wacc[g * 2] += (double)s0; wacc[g * 2 + 1] += (double)q0;
  }
}

}  // namespace bevgen
