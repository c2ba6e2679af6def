// Nearest-code vector quantisation (VectorQuantizer2.forward, modules/stage1/quantize.py:276-285):
//   d[r][j] = (|z_r|^2 + |e_j|^2) - 2 * <z_r, e_j>   (fp32, same association as the reference expression)
//   idx[r]  = argmin_j d[r][j]   (lowest index on ties);  z_q[r] = e[idx[r]]
// fp32 CUDA-core register-tiled product (bit-exact integer output is the contract, so no reduced-precision
// tensor-core operands here): CTA = 64 rows x 128 codes per sweep, 256 threads as 16x16, 4 rows x 8 codes per
// thread, running (min, argmin) kept per row and merged across the 16 lanes sharing a row with warp shuffles.
#include "common.cuh"
#include "kernels.cuh"

namespace bevgen {

constexpr int VQ_ROWS = 64, VQ_CODES = 128, VQ_KC = 32;

__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int D) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) {
    float v = x[(size_t)row * D + c];
    s += v * v;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

__global__ void __launch_bounds__(256) vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                         const float* __restrict__ zz, const float* __restrict__ ee,
                                                         long long* __restrict__ idx_out, float* __restrict__ zq_out, int rows, int n_codes,
                                                         int D) {
  __shared__ float As[VQ_KC][VQ_ROWS + 4];
  __shared__ float Bs[VQ_KC][VQ_CODES + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> codes, ty -> rows
  const int r0 = blockIdx.x * VQ_ROWS;
  float best[4];
  int besti[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = INFINITY; besti[i] = 0; }

  for (int c0 = 0; c0 < n_codes; c0 += VQ_CODES) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += VQ_KC) {
      // stage z tile (64 x 32) and codebook tile (128 x 32), transposed to [k][row] for conflict-free reads
      for (int i = threadIdx.x; i < VQ_ROWS * VQ_KC; i += 256) {
        const int r = i / VQ_KC, k = i % VQ_KC;
        As[k][r] = (r0 + r < rows) ? z[(size_t)(r0 + r) * D + k0 + k] : 0.f;
      }
      for (int i = threadIdx.x; i < VQ_CODES * VQ_KC; i += 256) {
        const int c = i / VQ_KC, k = i % VQ_KC;
        Bs[k][c] = (c0 + c < n_codes) ? cb[(size_t)(c0 + c) * D + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < VQ_KC; ++k) {
        float a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty * 4 + i;
      const float zr = (r < rows) ? zz[r] : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + tx + 16 * j;
        if (c < n_codes) {
          const float d = (zr + ee[c]) - 2.0f * acc[i][j];
          if (d < best[i] || (d == best[i] && c < besti[i])) { best[i] = d; besti[i] = c; }
        }
      }
    }
  }
  // merge across the 16 lanes (tx) that share each row; lanes of one row are contiguous within a half-warp
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    for (int o = 8; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best[i], o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti[i], o);
      if (ob < best[i] || (ob == best[i] && oi < besti[i])) { best[i] = ob; besti[i] = oi; }
    }
    const int r = r0 + ty * 4 + i;
    if (tx == 0 && r < rows) idx_out[r] = besti[i];
    if (zq_out != nullptr && r < rows) {
      for (int c = tx; c < D; c += 16) zq_out[(size_t)r * D + c] = cb[(size_t)besti[i] * D + c];
    }
  }
}

int launch_row_sqnorm(const float* x, float* out, int rows, int D, cudaStream_t st) {
  row_sqnorm_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, out, rows, D);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

int launch_vq_nearest(const float* z, const float* cb, const float* zz, const float* ee, long long* idx, float* zq, int rows, int n_codes,
                      int D, cudaStream_t st) {
  if (D % VQ_KC != 0 || rows <= 0) return BEVGEN_ERR_ARG;
  vq_nearest_kernel<<<(rows + VQ_ROWS - 1) / VQ_ROWS, 256, 0, st>>>(z, cb, zz, ee, idx, zq, rows, n_codes, D);
  return cudaGetLastError() == cudaSuccess ? BEVGEN_OK : BEVGEN_ERR_CUDA;
}

}  // namespace bevgen
