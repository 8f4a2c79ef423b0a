// eigh + soft pseudo-inverse of the SR / MinSR solve (quantax/optimizer/solver.py:94-101,142-146).
// The eigendecomposition is cuSOLVER syevd (a library call, reported separately in the bench);
// the pseudo-inverse epilogue is three small kernels on the eigenvector matrix.
#include <cusolverDn.h>
#include <stdlib.h>

#include "common.cuh"

namespace qtx {

static thread_local cusolverDnHandle_t g_solver = nullptr;

int solver_handle(cusolverDnHandle_t* h) {  // shared with pinv_rational.cu
  if (!g_solver) {
    cusolverStatus_t s = cusolverDnCreate(&g_solver);
    if (s != CUSOLVER_STATUS_SUCCESS) {
      set_error("cusolverDnCreate failed with status %d", (int)s);
      g_solver = nullptr;
      return QTX_ERR_SOLVER;
    }
  }
  *h = g_solver;
  return QTX_OK;
}

// rho[k] = sum_i Ut[k, i] b[i]   (Ut row-major = eigenvectors as rows), CTA per k
__global__ void __launch_bounds__(256) rows_dot_kernel(const double* __restrict__ Ut, int64_t n,
                                                       const double* __restrict__ b, double* __restrict__ rho) {
  __shared__ double red[8];
  const int64_t k = blockIdx.x;
  const double* row = Ut + k * n;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += row[i] * b[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    rho[k] = t;
  }
}

// coef[k] = lambda_k^+ * rho[k];  lambda^+ = 1/(lambda (1 + (tol/|lambda|)^6)), 0 where lambda == 0
__global__ void __launch_bounds__(1024) pinv_coef_kernel(const double* __restrict__ evals, int64_t n, double rtol,
                                                         double atol, double* __restrict__ rho) {
  __shared__ double red[32];
  __shared__ double bc;
  double mx = 0.0;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) mx = fmax(mx, fabs(evals[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    if (threadIdx.x == 0) bc = mx;
  }
  __syncthreads();
  const double tol = rtol * bc + atol;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
    double v = evals[k], a = fabs(v);
    double q = tol / a;
    double q2 = q * q;
    double f = 1.0 + q2 * q2 * q2;
    double inv = 1.0 / (v * f);
    rho[k] = (a > 0.0) ? inv * rho[k] : 0.0;
  }
}

// y[i] = sum_k Ut[k, i] coef[k]: thread per column i, rows split over grid.y
__global__ void __launch_bounds__(256) cols_comb_kernel(const double* __restrict__ Ut, int64_t n,
                                                        const double* __restrict__ coef, double* __restrict__ y) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t chunk = (n + gridDim.y - 1) / gridDim.y;
  int64_t k0 = blockIdx.y * chunk, k1 = k0 + chunk < n ? k0 + chunk : n;
  double acc = 0.0;
  for (int64_t k = k0; k < k1; ++k) acc += Ut[k * n + i] * coef[k];
  atomicAdd(y + i, acc);
}

// rho[k] = sum_without_noise_i(M[k, i] b[i])  (solver.py:114-125): the plain sum x, damped by
// 1 / (1 + (tol_snr / snr)^6) with snr = |mean| / sqrt(mean_i |r_i - mean|^2 / n).  CTA per row k; the
// second pass re-reads the row (L2-resident) so that the variance is formed from centred terms like the reference.
__global__ void __launch_bounds__(256) rows_dot_snr_kernel(const double* __restrict__ M, int64_t n, int64_t ld,
                                                           const double* __restrict__ b, double tol_snr,
                                                           double* __restrict__ rho) {
  __shared__ double red[8];
  __shared__ double bc;
  const int64_t k = blockIdx.x;
  const double* row = M + k * ld;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += row[i] * b[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    bc = t;
  }
  __syncthreads();
  const double x = bc, mean = x / (double)n;
  double var = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = row[i] * b[i] - mean;
    var += d * d;
  }
  var = warp_sum(var);
  __syncthreads();  // red is reused
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = var;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    const double sd = sqrt(t / (double)n / (double)n);
    const double snr = fabs(mean) / sd;
    const double q = tol_snr / snr, q2 = q * q;
    rho[k] = tol_snr > 1e-6 ? x / (1.0 + q2 * q2 * q2) : x;  // cond(tol_snr > 1e-6, ...) of solver.py:124
  }
}

// T += (rshift * trace(T) + ashift) I   (minnorm_shift_eig / lstsq_shift_eig, solver.py:50-77); one CTA
__global__ void __launch_bounds__(1024) trace_shift_kernel(double* __restrict__ T, int64_t n, double rshift,
                                                           double ashift) {
  __shared__ double red[32];
  __shared__ double bc;
  double tr = 0.0;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) tr += T[k * n + k];
  tr = warp_sum(tr);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tr;
  __syncthreads();
  if (threadIdx.x < 32) {
    tr = warp_sum(red[threadIdx.x]);
    if (threadIdx.x == 0) bc = rshift * tr + ashift;
  }
  __syncthreads();
  const double shift = bc;
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) T[k * n + k] += shift;
}

}  // namespace qtx

using namespace qtx;

static int syevd_lwork(int64_t n, int* lwork) {
  cusolverDnHandle_t h;
  int rc = solver_handle(&h);
  if (rc) return rc;
  cusolverStatus_t s =
      cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, nullptr, (int)n,
                                  nullptr, lwork);
  if (s != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolverDnDsyevd_bufferSize failed with status %d", (int)s);
    return QTX_ERR_SOLVER;
  }
  return QTX_OK;
}

extern "C" size_t qtx_pinv_eig_workspace_size(int64_t n) {
  int lwork = 0;
  if (n <= 0 || n > 46340 || syevd_lwork(n, &lwork)) return 0;
  // [lwork doubles | evals n | rho n]
  return ((size_t)lwork + 2 * (size_t)n) * sizeof(double) + 8192;
}

// eigh(T) [+ rho = U^T b (optionally SNR-damped) + y = U (lambda^+ o rho) when b != nullptr]
static int pinv_eig_impl(double* T, int64_t n, const double* b, double rtol, double atol, double tol_snr,
                         double* evals_out, double* y_out, int32_t* info_out, void* workspace, size_t workspace_bytes,
                         qtx_stream_t stream) {
  QTX_REQUIRE(T && info_out && workspace && n > 0 && n <= 46340 && (!b || y_out) && (b || evals_out), QTX_ERR_INVALID,
              "qtx_pinv_eig_solve / qtx_eigh: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cusolverDnHandle_t h;
  int rc = solver_handle(&h);
  if (rc) return rc;
  int lwork = 0;
  rc = syevd_lwork(n, &lwork);
  if (rc) return rc;
  QTX_REQUIRE(workspace_bytes >= ((size_t)lwork + 2 * (size_t)n + 512) * sizeof(double), QTX_ERR_INVALID,
              "qtx_pinv_eig_solve: workspace too small");
  double* work = (double*)workspace;
  double* evals = work + lwork + 512;  // 4 KB slack: the 64-bit API asks for slightly more than the legacy one
  double* rho = evals + n;
  if (rtol < 0) rtol = 1e-12;  // solver.py:12-21 for float64
  cusolverStatus_t s = cusolverDnSetStream(h, st);
  QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnSetStream failed (%d)", (int)s);
  // row-major symmetric == column-major symmetric; on exit T holds column-major eigenvectors,
  // i.e. row k of the buffer is eigenvector k.
  static const int algo = getenv("QTX_EIGH_ALGO") ? atoi(getenv("QTX_EIGH_ALGO")) : 0;
  if (algo == 1) {
    // experiment: 64-bit generic API (needs host + device workspaces)
    static thread_local cusolverDnParams_t params = nullptr;
    if (!params) cusolverDnCreateParams(&params);
    size_t dbytes = 0, hbytes = 0;
    s = cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, T, n,
                                    CUDA_R_64F, evals, CUDA_R_64F, &dbytes, &hbytes);
    QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "Xsyevd_bufferSize failed (%d)", (int)s);
    QTX_REQUIRE(dbytes <= (size_t)lwork * sizeof(double) + 4096, QTX_ERR_SOLVER, "Xsyevd needs %zu B > %zu B", dbytes,
                (size_t)lwork * sizeof(double));
    static thread_local void* hbuf = nullptr;
    static thread_local size_t hcap = 0;
    if (hbytes > hcap) { free(hbuf); hbuf = malloc(hbytes); hcap = hbytes; }
    s = cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, T, n, CUDA_R_64F,
                         evals, CUDA_R_64F, work, dbytes, hbuf, hbytes, info_out);
    QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnXsyevd failed (%d)", (int)s);
  } else {
    s = cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, T, (int)n, evals, work, lwork,
                         info_out);
    QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnDsyevd failed (%d)", (int)s);
  }
  count_launch();
  if (evals_out) QTX_CUDA(cudaMemcpyAsync(evals_out, evals, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (!b) return QTX_OK;  // qtx_eigh: decomposition only
  if (tol_snr > 1e-6) rows_dot_snr_kernel<<<(unsigned)n, 256, 0, st>>>(T, n, n, b, tol_snr, rho);
  else rows_dot_kernel<<<(unsigned)n, 256, 0, st>>>(T, n, b, rho);
  QTX_LAUNCH_CHECK();
  pinv_coef_kernel<<<1, 1024, 0, st>>>(evals, n, rtol, atol, rho);
  QTX_LAUNCH_CHECK();
  QTX_CUDA(cudaMemsetAsync(y_out, 0, n * sizeof(double), st));
  unsigned gx = (unsigned)((n + 255) / 256);
  int64_t split = (4ll * num_sms() + gx - 1) / gx;
  if (split > n) split = n;
  if (split < 1) split = 1;
  cols_comb_kernel<<<dim3(gx, (unsigned)split), 256, 0, st>>>(T, n, rho, y_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_pinv_eig_solve(double* T, int64_t n, const double* b, double rtol, double atol, double* evals_out,
                                  double* y_out, int32_t* info_out, void* workspace, size_t workspace_bytes,
                                  qtx_stream_t stream) {
  QTX_REQUIRE(b && y_out, QTX_ERR_INVALID, "qtx_pinv_eig_solve: bad argument");
  return pinv_eig_impl(T, n, b, rtol, atol, 0.0, evals_out, y_out, info_out, workspace, workspace_bytes, stream);
}

extern "C" int qtx_pinv_eig_solve_snr(double* T, int64_t n, const double* b, double rtol, double atol, double tol_snr,
                                      double* evals_out, double* y_out, int32_t* info_out, void* workspace,
                                      size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(b && y_out && tol_snr >= 0.0, QTX_ERR_INVALID, "qtx_pinv_eig_solve_snr: bad argument");
  return pinv_eig_impl(T, n, b, rtol, atol, tol_snr, evals_out, y_out, info_out, workspace, workspace_bytes, stream);
}

extern "C" int qtx_eigh(double* T, int64_t n, double* evals_out, int32_t* info_out, void* workspace,
                        size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(evals_out, QTX_ERR_INVALID, "qtx_eigh: bad argument");
  return pinv_eig_impl(T, n, nullptr, 0.0, 0.0, 0.0, evals_out, nullptr, info_out, workspace, workspace_bytes, stream);
}

extern "C" int qtx_rows_dot_snr(const double* M, int64_t nrows, int64_t n, int64_t ld, const double* b, double tol_snr,
                                double* rho_out, qtx_stream_t stream) {
  QTX_REQUIRE(M && b && rho_out && nrows > 0 && n > 0 && ld >= n && tol_snr >= 0.0, QTX_ERR_INVALID,
              "qtx_rows_dot_snr: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  rows_dot_snr_kernel<<<(unsigned)nrows, 256, 0, st>>>(M, n, ld, b, tol_snr, rho_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_pinv_apply(const double* Ut, int64_t n, const double* evals, double* rho_inout, double rtol,
                              double atol, double* y_out, qtx_stream_t stream) {
  QTX_REQUIRE(Ut && evals && rho_inout && y_out && n > 0, QTX_ERR_INVALID, "qtx_pinv_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (rtol < 0) rtol = 1e-12;
  pinv_coef_kernel<<<1, 1024, 0, st>>>(evals, n, rtol, atol, rho_inout);
  QTX_LAUNCH_CHECK();
  QTX_CUDA(cudaMemsetAsync(y_out, 0, n * sizeof(double), st));
  unsigned gx = (unsigned)((n + 255) / 256);
  int64_t split = (4ll * num_sms() + gx - 1) / gx;
  if (split > n) split = n;
  if (split < 1) split = 1;
  cols_comb_kernel<<<dim3(gx, (unsigned)split), 256, 0, st>>>(Ut, n, rho_inout, y_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

// ---- diagonal-shift solvers: (T + shift I)^-1 b by Cholesky (cuSOLVER potrf / potrs, library calls) ----------
static int potrf_lwork(int64_t n, int* lwork) {
  cusolverDnHandle_t h;
  int rc = solver_handle(&h);
  if (rc) return rc;
  cusolverStatus_t s = cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, (int)n, nullptr, (int)n, lwork);
  if (s != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolverDnDpotrf_bufferSize failed with status %d", (int)s);
    return QTX_ERR_SOLVER;
  }
  return QTX_OK;
}

extern "C" size_t qtx_shift_chol_workspace_size(int64_t n) {
  int lwork = 0;
  if (n <= 0 || n > 46340 || potrf_lwork(n, &lwork)) return 0;
  return (size_t)lwork * sizeof(double) + 4096;  // [potrf work | 4 KB: potrs info]
}

extern "C" int qtx_shift_chol_solve(double* T, int64_t n, const double* b, double rshift, double ashift, double* y_out,
                                    int32_t* info_out, void* workspace, size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(T && b && y_out && info_out && workspace && n > 0 && n <= 46340, QTX_ERR_INVALID,
              "qtx_shift_chol_solve: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cusolverDnHandle_t h;
  int rc = solver_handle(&h);
  if (rc) return rc;
  int lwork = 0;
  rc = potrf_lwork(n, &lwork);
  if (rc) return rc;
  QTX_REQUIRE(workspace_bytes >= (size_t)lwork * sizeof(double) + 4096, QTX_ERR_INVALID,
              "qtx_shift_chol_solve: workspace too small");
  double* work = (double*)workspace;
  int32_t* info2 = (int32_t*)((char*)workspace + (((size_t)lwork * sizeof(double) + 255) & ~(size_t)255));
  if (rshift < 0) rshift = 1e-12;  // solver.py:12-21 for float64
  cusolverStatus_t s = cusolverDnSetStream(h, st);
  QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnSetStream failed (%d)", (int)s);
  trace_shift_kernel<<<1, 1024, 0, st>>>(T, n, rshift, ashift);
  QTX_LAUNCH_CHECK();
  // row-major symmetric == column-major symmetric; the lower triangle is factorised in place
  s = cusolverDnDpotrf(h, CUBLAS_FILL_MODE_LOWER, (int)n, T, (int)n, work, lwork, info_out);
  QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnDpotrf failed (%d)", (int)s);
  QTX_CUDA(cudaMemcpyAsync(y_out, b, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  s = cusolverDnDpotrs(h, CUBLAS_FILL_MODE_LOWER, (int)n, 1, T, (int)n, y_out, (int)n, info2);
  QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnDpotrs failed (%d)", (int)s);
  count_launch(2);
  return QTX_OK;
}
