// Shared helpers for libqtx_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "qtx_b200.h"

namespace qtx {

constexpr unsigned FULL = 0xffffffffu;

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define QTX_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      qtx::set_error(__VA_ARGS__);     \
      return (code);                   \
    }                                  \
  } while (0)

#define QTX_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      qtx::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return QTX_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define QTX_LAUNCH_CHECK()                                                                   \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      qtx::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return QTX_ERR_CUDA;                                                                  \
    }                                                                                       \
    qtx::count_launch();                                                                    \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// log|cosh x| = |x| + log1p(exp(-2|x|)) - ln 2, stable for all x (no overflow at |x| > 89).
// float: two MUFU ops (ex2, lg2); the argument of lg2 lies in (1, 2] so its absolute error is
// ~2^-22, far below the rounding of the sum it enters.
__device__ __forceinline__ float lncosh(float x) {
  float ax = fabsf(x);
  float e = __expf(-2.0f * ax);
  return ax + (__logf(1.0f + e) - 0.69314718055994531f);
}
__device__ __forceinline__ double lncosh(double x) {
  double ax = fabs(x);
  return ax + (log1p(exp(-2.0 * ax)) - 0.69314718055994531);
}

// ---- Philox4x32-10: counter (c0..c3), key (k0,k1) -----------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = QTX_F32; };
template <> struct dtype_of<double> { static constexpr int value = QTX_F64; };

}  // namespace qtx
