// State-level symmetry projection (quantax/state/variational.py:262-266, symmetry.py:325-392):
//   images   s_g = s[perm_g]  (and -s[perm_g] for the Z2 block)           -> qtx_symm_images
//   combine  psi = sum_g w_g psi_g in container arithmetic (signed LSE)   -> qtx_symm_combine
//   jacobian O(s) = sum_g (w_g psi_g / psi) O(s_g)                        -> qtx_weighted_rowsum
// Byte gather / small reductions / one streaming pass: all HBM bound.
#include "common.cuh"

namespace qtx {

__global__ void symm_images_kernel(const int8_t* __restrict__ spins, int64_t ns, int N, const int32_t* __restrict__ perm,
                                   int nperm, int z2, int8_t* __restrict__ out) {
  const int nsymm = z2 ? 2 * nperm : nperm;
  const int64_t total = ns * nsymm * (int64_t)N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % N);
    const int64_t sg = e / N;
    const int g = (int)(sg % nsymm);
    const int64_t s = sg / nsymm;
    const int gp = g < nperm ? g : g - nperm;
    int8_t v = spins[s * N + perm[(size_t)gp * N + j]];
    out[e] = g < nperm ? v : (int8_t)-v;
  }
}

// kind 0: LogArray out (sign(b), emax + log|b|); kind 1: ScaleArray out (b, emax [+ log cmax folded by caller])
__global__ void symm_combine_kernel(const double* __restrict__ mult, const double* __restrict__ expo, int64_t ns,
                                    int nsymm, const double* __restrict__ w, int kind, double* __restrict__ mult_out,
                                    double* __restrict__ expo_out, double* __restrict__ coef_out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= ns) return;
  const double* m = mult + s * nsymm;
  const double* e = expo + s * nsymm;
  double emax = -INFINITY;
  for (int g = lane; g < nsymm; g += 32) emax = fmax(emax, e[g]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(FULL, emax, o));
  double b = 0.0;
  for (int g = lane; g < nsymm; g += 32) {
    const double r = (e[g] != emax) ? exp(e[g] - emax) : 1.0;  // utils/big_array.py:52-54
    b += m[g] * w[g] * r;
  }
  b = warp_sum(b);
  if (coef_out)  // w_g psi_g / psi: the weights of the projected log-derivative (variational.py:460-466)
    for (int g = lane; g < nsymm; g += 32) {
      const double r = (e[g] != emax) ? exp(e[g] - emax) : 1.0;
      coef_out[s * nsymm + g] = m[g] * w[g] * r / b;
    }
  if (lane == 0) {
    if (kind == 0) {
      mult_out[s] = (b > 0.0) ? 1.0 : ((b < 0.0) ? -1.0 : 0.0);
      expo_out[s] = emax + log(fabs(b));
    } else {
      mult_out[s] = b;
      expo_out[s] = emax;
    }
  }
}

// out[s, k] = sum_g coef[s, g] * J[s*nsymm + g, k]
template <typename T>
__global__ void __launch_bounds__(256) weighted_rowsum_kernel(const T* __restrict__ J, int64_t ldj,
                                                              const double* __restrict__ coef, int nsymm, int64_t np,
                                                              T* __restrict__ out, int64_t ldo) {
  const int64_t s = blockIdx.y;
  const double* c = coef + s * nsymm;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int g = 0; g < nsymm; ++g) acc += c[g] * (double)J[(s * nsymm + g) * ldj + k];
    out[s * ldo + k] = (T)acc;
  }
}

}  // namespace qtx

using namespace qtx;

extern "C" int qtx_symm_images(const int8_t* spins, int64_t ns, int N, const int32_t* perm, int nperm, int z2,
                               int8_t* out, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(spins && perm && out && N > 0 && nperm > 0, QTX_ERR_INVALID, "qtx_symm_images: bad argument");
  const int64_t total = ns * (z2 ? 2 : 1) * (int64_t)nperm * N;
  unsigned g = (unsigned)((total + 255) / 256);
  if (g > 16u * num_sms()) g = 16u * num_sms();
  symm_images_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(spins, ns, N, perm, nperm, z2, out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_symm_combine(const double* mult, const double* expo, int64_t ns, int nsymm, const double* weights,
                                int kind, double* mult_out, double* expo_out, double* coef_out, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(mult && expo && weights && mult_out && expo_out && nsymm > 0 && (kind == 0 || kind == 1), QTX_ERR_INVALID,
              "qtx_symm_combine: bad argument");
  symm_combine_kernel<<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(mult, expo, ns, nsymm, weights, kind,
                                                                                mult_out, expo_out, coef_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_weighted_rowsum(int dtype, const void* J, int64_t ldj, const double* coef, int64_t ns, int nsymm,
                                   int64_t np, void* out, int64_t ldo, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(J && coef && out && nsymm > 0 && np > 0 && ldj >= np && ldo >= np && ns <= 65535, QTX_ERR_INVALID,
              "qtx_weighted_rowsum: bad argument");
  unsigned gx = (unsigned)((np + 255) / 256);
  if (gx > 256) gx = 256;
  dim3 grid(gx, (unsigned)ns);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTX_F64) weighted_rowsum_kernel<double><<<grid, 256, 0, st>>>((const double*)J, ldj, coef, nsymm, np, (double*)out, ldo);
  else if (dtype == QTX_F32) weighted_rowsum_kernel<float><<<grid, 256, 0, st>>>((const float*)J, ldj, coef, nsymm, np, (float*)out, ldo);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_weighted_rowsum: bad dtype %d", dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
