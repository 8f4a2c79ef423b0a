// T = A A^T with plain FP64 FMA on CUDA cores (nslices = -1).  This is the development /
// cross-check path for the tensor-core Gram in gram_tc.cu; the product path is the tcgen05 kernel.
#include "common.cuh"

namespace qtx {

template <typename T>
__global__ void __launch_bounds__(256) gram_fma_kernel(const T* __restrict__ A, int64_t ns, int64_t np, int64_t ld,
                                                       double* __restrict__ Tout, int accum) {
  // lower-triangular tile (bi >= bj) from a linear block index
  int64_t b = blockIdx.x;
  int64_t bi = (int64_t)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
  while (bi * (bi + 1) / 2 > b) --bi;
  int64_t bj = b - bi * (bi + 1) / 2;
  __shared__ double As[16][65];
  __shared__ double Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
  double acc[4][4] = {};
  const int64_t ra = bi * 64 + lr, rb = bj * 64 + lr;
  for (int64_t k0 = 0; k0 < np; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int64_t k = k0 + lk + e;
      As[lk + e][lr] = (ra < ns && k < np) ? (double)A[ra * ld + k] : 0.0;
      Bs[lk + e][lr] = (rb < ns && k < np) ? (double)A[rb * ld + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[4], bb[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a[e] = As[k][ty * 4 + e];
        bb[e] = Bs[k][tx * 4 + e];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t r = bi * 64 + ty * 4 + i, c = bj * 64 + tx * 4 + j;
      if (r < ns && c < ns) {
        double v = acc[i][j];
        if (bi != bj || c <= r) {
          Tout[r * ns + c] = accum ? Tout[r * ns + c] + v : v;
          if (r != c) Tout[c * ns + r] = accum ? Tout[c * ns + r] + v : v;
        }
      }
    }
}

int gram_fma(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, double* Tout, int accum,
             cudaStream_t st) {
  int64_t nb = (ns + 63) / 64;
  int64_t nblk = nb * (nb + 1) / 2;
  if (dtype == QTX_F64)
    gram_fma_kernel<double><<<(unsigned)nblk, 256, 0, st>>>((const double*)A, ns, np, ld, Tout, accum);
  else
    gram_fma_kernel<float><<<(unsigned)nblk, 256, 0, st>>>((const float*)A, ns, np, ld, Tout, accum);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

}  // namespace qtx
