// HBM-bound dense helpers of the SR/MinSR step: column means, centring/scaling, Ebar,
// A^T y and A x products, parameter update.  All are single passes over the [ns, np] Jacobian with
// coalesced (vectorised where aligned) accesses; reductions over samples are split over the grid
// and combined with float64 atomics.
#include "common.cuh"

namespace qtx {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// out[k] (+)= alpha * sum_s w[s] A[s,k]; thread owns 2 adjacent columns, grid.y splits the rows
template <typename T, bool SQ = false>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ A, int64_t ns, int64_t np, int64_t ld,
                                                     const double* __restrict__ w, double alpha,
                                                     double* __restrict__ out, bool vec) {
  int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (k >= np) return;
  int64_t chunk = (ns + gridDim.y - 1) / gridDim.y;
  int64_t s0 = blockIdx.y * chunk, s1 = s0 + chunk < ns ? s0 + chunk : ns;
  double a0 = 0.0, a1 = 0.0;
  const bool two = k + 1 < np;
  if (vec && two) {
    using V = typename Vec2<T>::type;
#pragma unroll 4
    for (int64_t s = s0; s < s1; ++s) {
      V v = *reinterpret_cast<const V*>(A + s * ld + k);
      double ws = w ? w[s] : 1.0;
      double x0 = (double)v.x, x1 = (double)v.y;
      if (SQ) { x0 *= x0; x1 *= x1; }
      a0 += ws * x0;
      a1 += ws * x1;
    }
  } else {
    for (int64_t s = s0; s < s1; ++s) {
      double ws = w ? w[s] : 1.0;
      double x0 = (double)A[s * ld + k], x1 = two ? (double)A[s * ld + k + 1] : 0.0;
      if (SQ) { x0 *= x0; x1 *= x1; }
      a0 += ws * x0;
      if (two) a1 += ws * x1;
    }
  }
  atomicAdd(out + k, alpha * a0);
  if (two) atomicAdd(out + k + 1, alpha * a1);
}

template <typename T>
__global__ void __launch_bounds__(256) center_scale_kernel(T* __restrict__ A, int64_t ns, int64_t np, int64_t ld,
                                                           const double* __restrict__ mean,
                                                           const double* __restrict__ scale) {
  const int64_t s = blockIdx.y;
  const double sc = scale ? scale[s] : 1.0;
  T* row = A + s * ld;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    double v = (double)row[k];
    if (mean) v -= mean[k];
    row[k] = (T)(v * sc);
  }
}

// v[s] = sum_k A[s,k] x[k]  (CTA per row)
template <typename T>
__global__ void __launch_bounds__(256) matvec_kernel(const T* __restrict__ A, int64_t np, int64_t ld,
                                                     const double* __restrict__ x, double* __restrict__ v) {
  __shared__ double red[8];
  const int64_t s = blockIdx.x;
  const T* row = A + s * ld;
  double acc = 0.0;
  for (int64_t k = threadIdx.x; k < np; k += blockDim.x) acc += (double)row[k] * x[k];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    v[s] = t;
  }
}

// single-CTA statistics of the local energies (SR.get_Ebar)
__global__ void __launch_bounds__(1024) ebar_kernel(const double* __restrict__ eloc, const double* __restrict__ rw,
                                                    int64_t ns, double* __restrict__ ebar,
                                                    double* __restrict__ stats) {
  __shared__ double red[3][32];
  __shared__ double bc[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a = 0.0, b = 0.0;  // sum Eloc*rw, sum Eloc
  for (int64_t s = tid; s < ns; s += blockDim.x) {
    double e = eloc[s], r = rw ? rw[s] : 1.0;
    a += e * r;
    b += e;
  }
  a = warp_sum(a); b = warp_sum(b);
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
  __syncthreads();
  if (warp == 0) {
    a = lane < (blockDim.x >> 5) ? red[0][lane] : 0.0;
    b = lane < (blockDim.x >> 5) ? red[1][lane] : 0.0;
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { bc[0] = a / (double)ns; bc[1] = b / (double)ns; }
  }
  __syncthreads();
  const double emean = bc[0], emean_plain = bc[1];
  double v = 0.0;
  for (int64_t s = tid; s < ns; s += blockDim.x) {
    double e = eloc[s], r = rw ? rw[s] : 1.0;
    double d = e - emean;
    v += d * d * r;
    if (ebar) ebar[s] = (e - emean_plain) * sqrt(r / (double)ns);
  }
  v = warp_sum(v);
  if (lane == 0) red[2][warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? red[2][lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0 && stats) { stats[0] = emean; stats[1] = v / (double)ns; }
  }
}

__global__ void finite_flag_kernel(const double* __restrict__ step, int64_t np, int32_t* __restrict__ flag) {
  int bad = 0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x)
    bad |= !isfinite(step[k]);
  if (__any_sync(FULL, bad) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}

template <typename T>
__global__ void apply_update_kernel(T* __restrict__ params, const double* __restrict__ step, double lr, int64_t np,
                                    const int32_t* __restrict__ flag) {
  if (*flag == 0) return;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x)
    params[k] = params[k] + (T)(-(step[k] * lr));
}


// S[i, j] += alpha * x[i] * x[j]   (TimeEvol: S - outer(conj(Omean), Omean), quantax/optimizer/time_evol.py:113)
__global__ void rank1_update_kernel(int64_t n, double alpha, const double* __restrict__ x, double* __restrict__ S) {
  const int64_t total = n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    S[e] += alpha * x[e / n] * x[e % n];
}

// ---- vector helpers of the momentum optimizers (SPRING / MARCH / AdamSR, quantax/optimizer/sr.py:198-429) ----
__global__ void axpby_kernel(int64_t n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    y[k] = a * x[k] + (b == 0.0 ? 0.0 : b * y[k]);
}
// out[k] = x[k] / d[k] + c * z[k]   (z nullable)
__global__ void div_add_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ d, double c,
                               const double* __restrict__ z, double* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    out[k] = x[k] / d[k] + (z ? c * z[k] : 0.0);
}
// V[k] = beta V[k] + (1 - beta) |x[k] - y[k]|^2   (y nullable)
__global__ void second_moment_kernel(int64_t n, double beta, const double* __restrict__ x, const double* __restrict__ y,
                                     double* __restrict__ V) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const double d = x[k] - (y ? y[k] : 0.0);
    V[k] = beta * V[k] + (1.0 - beta) * d * d;
  }
}
// out[k] = (v[k] / corr)^(1/4) + eps
__global__ void fourth_root_kernel(int64_t n, const double* __restrict__ v, double corr, double eps,
                                   double* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    out[k] = sqrt(sqrt(v[k] / corr)) + eps;
}
template <typename T>
__global__ void __launch_bounds__(256) scale_columns_kernel(T* __restrict__ A, int64_t np, int64_t ld,
                                                            const double* __restrict__ d) {
  T* row = A + (int64_t)blockIdx.y * ld;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x)
    row[k] = (T)((double)row[k] / d[k]);
}

__global__ void set_i32_kernel(int32_t* p, int32_t v) { *p = v; }

template <typename T, bool SQ = false>
static int colsum_launch(const void* A, int64_t ns, int64_t np, int64_t ld, const double* w, double alpha,
                         double* out, bool accumulate, cudaStream_t st) {
  if (!accumulate) QTX_CUDA(cudaMemsetAsync(out, 0, np * sizeof(double), st));
  unsigned gx = (unsigned)(((np + 1) / 2 + 255) / 256);
  int64_t split = (8ll * num_sms() + gx - 1) / gx;
  if (split < 1) split = 1;
  if (split > ns) split = ns;
  if (split > 1024) split = 1024;
  bool vec = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) % (2 * sizeof(T))) == 0);
  colsum_kernel<T, SQ><<<dim3(gx, (unsigned)split), 256, 0, st>>>((const T*)A, ns, np, ld, w, alpha, out, vec);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

}  // namespace qtx

using namespace qtx;

extern "C" int qtx_colmean(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, const double* weight,
                           double* mean_out, qtx_stream_t stream) {
  QTX_REQUIRE(A && mean_out && ns > 0 && np > 0 && ld >= np, QTX_ERR_INVALID, "qtx_colmean: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTX_F64) return colsum_launch<double>(A, ns, np, ld, weight, 1.0 / (double)ns, mean_out, false, st);
  if (dtype == QTX_F32) return colsum_launch<float>(A, ns, np, ld, weight, 1.0 / (double)ns, mean_out, false, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_colmean: bad dtype %d", dtype);
}

extern "C" int qtx_col_sumsq(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, double* out,
                             qtx_stream_t stream) {
  QTX_REQUIRE(A && out && ns > 0 && np > 0 && ld >= np, QTX_ERR_INVALID, "qtx_col_sumsq: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTX_F64) return colsum_launch<double, true>(A, ns, np, ld, nullptr, 1.0, out, false, st);
  if (dtype == QTX_F32) return colsum_launch<float, true>(A, ns, np, ld, nullptr, 1.0, out, false, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_col_sumsq: bad dtype %d", dtype);
}

extern "C" int qtx_matvec_t(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, const double* y,
                            double* x_out, int accumulate, qtx_stream_t stream) {
  QTX_REQUIRE(A && y && x_out && ns > 0 && np > 0 && ld >= np, QTX_ERR_INVALID, "qtx_matvec_t: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTX_F64) return colsum_launch<double>(A, ns, np, ld, y, 1.0, x_out, accumulate != 0, st);
  if (dtype == QTX_F32) return colsum_launch<float>(A, ns, np, ld, y, 1.0, x_out, accumulate != 0, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_matvec_t: bad dtype %d", dtype);
}

extern "C" int qtx_center_scale(int dtype, void* A, int64_t ns, int64_t np, int64_t ld, const double* mean,
                                const double* scale, qtx_stream_t stream) {
  QTX_REQUIRE(A && ns > 0 && np > 0 && ld >= np && ns < 65536 * 32768ll, QTX_ERR_INVALID,
              "qtx_center_scale: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned gx = (unsigned)((np + 255) / 256);
  if (gx > 64) gx = 64;
  QTX_REQUIRE(ns <= 65535, QTX_ERR_UNSUPPORTED, "qtx_center_scale: ns > 65535 rows per call");
  dim3 grid(gx, (unsigned)ns);
  if (dtype == QTX_F64) center_scale_kernel<double><<<grid, 256, 0, st>>>((double*)A, ns, np, ld, mean, scale);
  else if (dtype == QTX_F32) center_scale_kernel<float><<<grid, 256, 0, st>>>((float*)A, ns, np, ld, mean, scale);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_center_scale: bad dtype %d", dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_matvec(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, const double* x,
                          double* v_out, qtx_stream_t stream) {
  QTX_REQUIRE(A && x && v_out && ns > 0 && np > 0 && ld >= np, QTX_ERR_INVALID, "qtx_matvec: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTX_F64) matvec_kernel<double><<<(unsigned)ns, 256, 0, st>>>((const double*)A, np, ld, x, v_out);
  else if (dtype == QTX_F32) matvec_kernel<float><<<(unsigned)ns, 256, 0, st>>>((const float*)A, np, ld, x, v_out);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_matvec: bad dtype %d", dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_ebar(const double* eloc, const double* rw, int64_t ns, double* ebar_out, double* stats_out,
                        qtx_stream_t stream) {
  QTX_REQUIRE(eloc && ns > 0, QTX_ERR_INVALID, "qtx_ebar: bad argument");
  ebar_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(eloc, rw, ns, ebar_out, stats_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_apply_update(int model_dtype, void* params, const double* step, double lr, int64_t np,
                                int32_t* flag_out, qtx_stream_t stream) {
  QTX_REQUIRE(params && step && flag_out && np > 0, QTX_ERR_INVALID, "qtx_apply_update: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  set_i32_kernel<<<1, 1, 0, st>>>(flag_out, 1);
  QTX_LAUNCH_CHECK();
  unsigned g = (unsigned)((np + 255) / 256);
  if (g > 4u * num_sms()) g = 4u * num_sms();
  finite_flag_kernel<<<g, 256, 0, st>>>(step, np, flag_out);
  QTX_LAUNCH_CHECK();
  if (model_dtype == QTX_F32) apply_update_kernel<float><<<g, 256, 0, st>>>((float*)params, step, lr, np, flag_out);
  else if (model_dtype == QTX_F64)
    apply_update_kernel<double><<<g, 256, 0, st>>>((double*)params, step, lr, np, flag_out);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_apply_update: bad dtype %d", model_dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

static unsigned vec_grid(int64_t n) {
  unsigned g = (unsigned)((n + 255) / 256);
  unsigned cap = 8u * (unsigned)num_sms();
  return g > cap ? cap : (g ? g : 1);
}

extern "C" int qtx_rank1_update(int64_t n, double alpha, const double* x, double* S, qtx_stream_t stream) {
  if (n == 0) return QTX_OK;
  QTX_REQUIRE(x && S && n > 0, QTX_ERR_INVALID, "qtx_rank1_update: bad argument");
  rank1_update_kernel<<<vec_grid(n * n), 256, 0, (cudaStream_t)stream>>>(n, alpha, x, S);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_axpby(int64_t n, double a, const double* x, double b, double* y, qtx_stream_t stream) {
  if (n == 0) return QTX_OK;
  QTX_REQUIRE(x && y && n > 0, QTX_ERR_INVALID, "qtx_axpby: bad argument");
  axpby_kernel<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>(n, a, x, b, y);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_div_add(int64_t n, const double* x, const double* d, double c, const double* z, double* out,
                           qtx_stream_t stream) {
  if (n == 0) return QTX_OK;
  QTX_REQUIRE(x && d && out && n > 0, QTX_ERR_INVALID, "qtx_div_add: bad argument");
  div_add_kernel<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>(n, x, d, c, z, out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_second_moment(int64_t n, double beta, const double* x, const double* y, double* V,
                                 qtx_stream_t stream) {
  if (n == 0) return QTX_OK;
  QTX_REQUIRE(x && V && n > 0, QTX_ERR_INVALID, "qtx_second_moment: bad argument");
  second_moment_kernel<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>(n, beta, x, y, V);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_fourth_root(int64_t n, const double* v, double corr, double eps, double* out, qtx_stream_t stream) {
  if (n == 0) return QTX_OK;
  QTX_REQUIRE(v && out && n > 0 && corr != 0.0, QTX_ERR_INVALID, "qtx_fourth_root: bad argument");
  fourth_root_kernel<<<vec_grid(n), 256, 0, (cudaStream_t)stream>>>(n, v, corr, eps, out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_scale_columns(int dtype, void* A, int64_t ns, int64_t np, int64_t ld, const double* d,
                                 qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(A && d && np > 0 && ld >= np && ns <= 65535, QTX_ERR_INVALID, "qtx_scale_columns: bad argument");
  unsigned gx = (unsigned)((np + 255) / 256);
  if (gx > 64) gx = 64;
  dim3 grid(gx, (unsigned)ns);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTX_F64) scale_columns_kernel<double><<<grid, 256, 0, st>>>((double*)A, np, ld, d);
  else if (dtype == QTX_F32) scale_columns_kernel<float><<<grid, 256, 0, st>>>((float*)A, np, ld, d);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_scale_columns: bad dtype %d", dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
