// RBM_Conv / SingleConv (quantax/model/shallow_nets.py:129-190): psi(s) = prod_{c,r} cosh(theta_{c,r}),
//   theta_{c,r} = b_c + sum_d K_c[d] * s[(r + d - lo) mod L]      (one full-lattice circular convolution:
//   eqx.nn.Conv(kernel_size = lattice extent, padding = "SAME", padding_mode = "CIRCULAR"); cross-correlation,
//   lo = (L - 1) / 2 per dimension).
// This is a dense RBM with M = C * N hidden units and tied weights W[(c, r), j] = K_c[(j - r + lo) mod L], so the
// sweep / Oloc / forward kernels of rbm.cu are reused on the expanded weights (qtx_rbm_conv_expand, 4 * C * N^2
// bytes, rebuilt whenever the parameters change) and only the log-derivative needs its own kernel:
//   O[s, c*N + d] = sum_r tanh(theta_{c,r}) * s[(r + d - lo) mod L],   O[s, C*N + c] = sum_r tanh(theta_{c,r}).
#include "common.cuh"

namespace qtx {

// position arithmetic on an Lx x Ly lattice (chains: Lx = 1): j = (r + d - lo) mod L per dimension
__device__ __forceinline__ int conv_src(int r, int d, int lx, int ly) {
  const int ry = r % ly, rx = r / ly, dy = d % ly, dx = d / ly;
  int x = rx + dx - (lx - 1) / 2, y = ry + dy - (ly - 1) / 2;
  x %= lx; if (x < 0) x += lx;
  y %= ly; if (y < 0) y += ly;
  return x * ly + y;
}

template <typename T>
__global__ void rbm_conv_expand_kernel(const T* __restrict__ K, const T* __restrict__ bias, int C, int lx, int ly,
                                       T* __restrict__ W, T* __restrict__ b) {
  const int N = lx * ly, M = C * N;
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(e % N), m = (int)(e / N);
    const int c = m / N, r = m - c * N;
    W[(int64_t)m * N + conv_src(r, d, lx, ly)] = K[c * N + d];
  }
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) b[m] = bias ? bias[m / N] : T(0);
}

// one CTA per sample; tanh(theta) [M] and the spins [N] are staged in shared memory
template <typename T, typename OutT>
__global__ void __launch_bounds__(256) rbm_conv_jacobian_kernel(const T* __restrict__ theta, const int8_t* __restrict__ spins,
                                                                int C, int lx, int ly, OutT* __restrict__ out, int64_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = lx * ly, M = C * N;
  T* th = reinterpret_cast<T*>(smem_raw);
  T* sp = th + M;
  const int64_t s = blockIdx.x;
  for (int m = threadIdx.x; m < M; m += blockDim.x) th[m] = tanh(theta[s * M + m]);
  for (int j = threadIdx.x; j < N; j += blockDim.x) sp[j] = (T)spins[s * N + j];
  __syncthreads();
  OutT* o = out + s * ld;
  for (int e = threadIdx.x; e < M; e += blockDim.x) {
    const int c = e / N, d = e - c * N;
    T acc = 0;
    for (int r = 0; r < N; ++r) acc += th[c * N + r] * sp[conv_src(r, d, lx, ly)];
    o[e] = (OutT)acc;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    T acc = 0;
    for (int r = 0; r < N; ++r) acc += th[c * N + r];
    o[M + c] = (OutT)acc;
  }
}

}  // namespace qtx

using namespace qtx;

extern "C" int qtx_rbm_conv_expand(int model_dtype, const void* kernel, const void* bias, int channels, int lx, int ly,
                                   void* W_out, void* b_out, qtx_stream_t stream) {
  QTX_REQUIRE(kernel && W_out && b_out && channels > 0 && lx > 0 && ly > 0, QTX_ERR_INVALID,
              "qtx_rbm_conv_expand: bad argument");
  const int64_t total = (int64_t)channels * lx * ly * lx * ly;
  unsigned g = (unsigned)((total + 255) / 256);
  if (g > 8u * num_sms()) g = 8u * num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32)
    rbm_conv_expand_kernel<float><<<g, 256, 0, st>>>((const float*)kernel, (const float*)bias, channels, lx, ly,
                                                     (float*)W_out, (float*)b_out);
  else if (model_dtype == QTX_F64)
    rbm_conv_expand_kernel<double><<<g, 256, 0, st>>>((const double*)kernel, (const double*)bias, channels, lx, ly,
                                                      (double*)W_out, (double*)b_out);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_conv_expand: bad dtype %d", model_dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_rbm_conv_jacobian(int model_dtype, const void* theta, const int8_t* spins, int64_t ns, int channels,
                                     int lx, int ly, int out_dtype, void* out, int64_t ld, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  const int N = lx * ly, M = channels * N;
  QTX_REQUIRE(theta && spins && out && channels > 0 && N > 0 && ld >= M + channels, QTX_ERR_INVALID,
              "qtx_rbm_conv_jacobian: bad argument");
  QTX_REQUIRE(out_dtype == QTX_F32 || out_dtype == QTX_F64, QTX_ERR_INVALID, "qtx_rbm_conv_jacobian: bad out dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = model_dtype == QTX_F64 ? 8 : 4;
  const size_t smem = (size_t)(M + N) * es;
  QTX_REQUIRE(smem <= 200 * 1024, QTX_ERR_UNSUPPORTED, "qtx_rbm_conv_jacobian: C * N too large for shared memory");
#define QTX_LAUNCH_RCJ(T, OutT)                                                                                   \
  do {                                                                                                            \
    auto k = rbm_conv_jacobian_kernel<T, OutT>;                                                                   \
    if (smem > 48 * 1024) QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k<<<(unsigned)ns, 256, smem, st>>>((const T*)theta, spins, channels, lx, ly, (OutT*)out, ld);                  \
  } while (0)
  if (model_dtype == QTX_F32 && out_dtype == QTX_F64) QTX_LAUNCH_RCJ(float, double);
  else if (model_dtype == QTX_F32) QTX_LAUNCH_RCJ(float, float);
  else if (model_dtype == QTX_F64 && out_dtype == QTX_F64) QTX_LAUNCH_RCJ(double, double);
  else if (model_dtype == QTX_F64) QTX_LAUNCH_RCJ(double, float);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_conv_jacobian: bad dtype %d", model_dtype);
#undef QTX_LAUNCH_RCJ
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
