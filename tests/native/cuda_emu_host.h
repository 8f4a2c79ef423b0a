// Host-side stubs for running the ENTRY POINTS of csrc/pinv_rational.cu on the CPU (TEST INFRASTRUCTURE ONLY): the
// error / launch macros of common.cuh, a stream type, and a stand-in for the four cuSOLVER calls with LAPACK
// semantics (column-major, partial pivoting, 1-based pivots).  The stand-in scribbles over the whole workspace it was
// promised, so an overlap in the caller's workspace layout corrupts the result and fails the test.
#pragma once
#include <complex>
#include <cstdarg>
#include <cstdio>

#include "qtx_b200.h"

typedef void* cudaStream_t;

#define QTX_LAUNCH(kernel, grid, block, stream, ...) emu_launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#define QTX_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
  emu_launch((grid), (block), [&] { kernel(__VA_ARGS__); }, (smem))

namespace qtx {
inline char g_emu_error[512] = "";
inline long g_emu_launches = 0;
inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_emu_error, sizeof(g_emu_error), fmt, ap);
  va_end(ap);
}
inline void count_launch(int n = 1) { g_emu_launches += n; }
inline int num_sms() { return 2; }  // small on purpose: grid-stride loops get exercised
}  // namespace qtx

#define QTX_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      qtx::set_error(__VA_ARGS__);   \
      return (code);                 \
    }                                \
  } while (0)
#define QTX_LAUNCH_CHECK() qtx::count_launch()

// ---- cuSOLVER stand-in -------------------------------------------------------------------------------------------
struct EmuSolver {
  int dummy;
};
typedef EmuSolver* cusolverDnHandle_t;
enum cusolverStatus_t { CUSOLVER_STATUS_SUCCESS = 0, CUSOLVER_STATUS_INVALID_VALUE = 3 };
enum cublasOperation_t { CUBLAS_OP_N = 0, CUBLAS_OP_T = 1, CUBLAS_OP_C = 2 };

inline cusolverStatus_t cusolverDnSetStream(cusolverDnHandle_t, cudaStream_t) { return CUSOLVER_STATUS_SUCCESS; }

inline cusolverStatus_t cusolverDnZgetrf_bufferSize(cusolverDnHandle_t, int m, int n, cuDoubleComplex*, int lda,
                                                    int* lwork) {
  if (m < 0 || n < 0 || lda < (m > 1 ? m : 1)) return CUSOLVER_STATUS_INVALID_VALUE;
  *lwork = 3 * n + 7;
  return CUSOLVER_STATUS_SUCCESS;
}

inline cusolverStatus_t cusolverDnZgetrf(cusolverDnHandle_t, int m, int n, cuDoubleComplex* A, int lda,
                                         cuDoubleComplex* work, int* ipiv, int* info) {
  typedef std::complex<double> cd;
  if (m != n || !A || !work || !ipiv || !info) return CUSOLVER_STATUS_INVALID_VALUE;
  for (int i = 0; i < 3 * n + 7; ++i) work[i] = {1e300, -1e300};  // claim the promised workspace
  cd* a = reinterpret_cast<cd*>(A);
  *info = 0;
  for (int k = 0; k < n; ++k) {  // column-major: a[i + j * lda]
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (std::abs(a[i + (size_t)k * lda]) > std::abs(a[p + (size_t)k * lda])) p = i;
    ipiv[k] = p + 1;
    if (a[p + (size_t)k * lda] == cd(0.0)) {
      if (*info == 0) *info = k + 1;
      continue;
    }
    if (p != k)
      for (int j = 0; j < n; ++j) std::swap(a[k + (size_t)j * lda], a[p + (size_t)j * lda]);
    for (int i = k + 1; i < n; ++i) a[i + (size_t)k * lda] /= a[k + (size_t)k * lda];
    for (int j = k + 1; j < n; ++j) {
      const cd u = a[k + (size_t)j * lda];
      for (int i = k + 1; i < n; ++i) a[i + (size_t)j * lda] -= a[i + (size_t)k * lda] * u;
    }
  }
  return CUSOLVER_STATUS_SUCCESS;
}

inline cusolverStatus_t cusolverDnZgetrs(cusolverDnHandle_t, cublasOperation_t trans, int n, int nrhs,
                                         const cuDoubleComplex* A, int lda, const int* ipiv, cuDoubleComplex* B,
                                         int ldb, int* info) {
  typedef std::complex<double> cd;
  if (trans != CUBLAS_OP_N || !A || !ipiv || !B || !info || ldb < n) return CUSOLVER_STATUS_INVALID_VALUE;
  const cd* a = reinterpret_cast<const cd*>(A);
  cd* b = reinterpret_cast<cd*>(B);
  *info = 0;
  for (int r = 0; r < nrhs; ++r) {
    cd* x = b + (size_t)r * ldb;
    for (int k = 0; k < n; ++k) std::swap(x[k], x[ipiv[k] - 1]);
    for (int k = 0; k < n; ++k)
      for (int i = k + 1; i < n; ++i) x[i] -= a[i + (size_t)k * lda] * x[k];
    for (int k = n - 1; k >= 0; --k) {
      x[k] /= a[k + (size_t)k * lda];
      for (int i = 0; i < k; ++i) x[i] -= a[i + (size_t)k * lda] * x[k];
    }
  }
  return CUSOLVER_STATUS_SUCCESS;
}
