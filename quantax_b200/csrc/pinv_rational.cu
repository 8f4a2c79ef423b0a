// Soft pseudo-inverse of the SR / MinSR solve WITHOUT an eigendecomposition
// (quantax/optimizer/solver.py:94-111,142-146 compute y = U f(lambda) U^T b after eigh).
//
//   f(lambda) = 1 / (lambda (1 + (c/|lambda|)^6)) = lambda^5 / (lambda^6 + c^6),   c = rtol max|lambda| + atol,
//
// and lambda^5 / P(lambda) with P = lambda^6 + c^6 is P'/(6P), so over the roots z_k = c exp(i pi (2k+1)/6) of P
//
//   f(T) b = (1/6) sum_{k<6} (T - z_k)^-1 b = (1/3) Re sum_{k=0,1,2} (T - z_k I)^-1 b        (T, b real)
//
// exactly: three complex-symmetric linear solves (cuSOLVER Zgetrf/Zgetrs, library calls) replace syevd, whose
// BLAS-2 tridiagonalisation is the largest single cost of the MinSR step.  Directions with |lambda| << c give terms
// of size 1/c that cancel between the three shifts, so each solve is refined with residuals evaluated in
// double-double arithmetic (the LU factors are reused: a few triangular solves), the solution is kept as a
// double-double vector and the three shifts are summed in double-double before the result is rounded once.
// Against an exact evaluation of f(T) b for the same float64 T this is MORE accurate than the eigenvalue route
// when eigenvalues lie near the cut-off (oracle/pinv_rational.py, tests/test_pinv_rational_cpu.py).
// max|lambda| comes from a Lanczos recurrence (three-term, no reorthogonalisation) and bisection on the Sturm count.
#ifdef QTX_HOST_EMULATION  // tests/native: this file runs unmodified on the CPU (CUDA threads = std::thread,
#include "cuda_emu.h"       // cuSOLVER = a LAPACK-style stand-in); see tests/test_pinv_rational_emu_cpu.py
#include "cuda_emu_host.h"
#else
#include <cuComplex.h>
#include <cusolverDn.h>

#include "common.cuh"
#define QTX_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif
#include "dd_math.cuh"
#include "zldlt.cuh"

namespace qtx {

int solver_handle(cusolverDnHandle_t* h);  // solver.cu (tests/native/pinv_rational_emu.cpp under emulation)

constexpr int kLanczosMaxSteps = 1024;
#ifdef QTX_HOST_EMULATION
constexpr unsigned kLanczosThreads = 128;  // fewer std::threads per launch; the kernels take any multiple of 32
constexpr unsigned kRowThreads = 64;
#else
constexpr unsigned kLanczosThreads = 1024;
constexpr unsigned kRowThreads = 256;  // dd_residual_kernel: threads per matrix row
#endif

// ---- double-double helpers: dd_math.cuh (shared with the CPU test of the same arithmetic) ----------------------
__device__ __forceinline__ dd dd_shfl_xor(dd a, int o) {
  return {__shfl_xor_sync(FULL, a.hi, o), __shfl_xor_sync(FULL, a.lo, o)};
}

// block-wide sum of one double per thread (blockDim.x <= 1024, a multiple of 32); every thread gets the result
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();  // red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// ---- Lanczos for max|lambda| -------------------------------------------------------------------------------
// deterministic start vector from an integer hash (oracle/pinv_rational.py start_vector), normalised
__global__ void __launch_bounds__(1024) lanczos_init_kernel(int64_t n, double* __restrict__ v,
                                                            double* __restrict__ vprev, double* __restrict__ state) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    uint64_t x = ((uint64_t)i + 1ull) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    const double u = (double)(x >> 11) * 0x1p-52 - 1.0;
    v[i] = u;
    vprev[i] = 0.0;
    acc += u * u;
  }
  const double inv = rsqrt(block_sum(acc, red));
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) v[i] *= inv;
  if (threadIdx.x == 0) {
    state[0] = 0.0;  // beta of the previous step
    state[1] = 0.0;  // running scale max(|alpha|, beta)
  }
}

// one step after w = T v:  alpha = w.v;  w -= alpha v + beta_prev v_prev;  second pass against v;  beta = |w|;
// (v_prev, v) <- (v, w / beta).  A breakdown (beta <= 1e-13 scale) zeroes the following vectors.
__global__ void __launch_bounds__(1024) lanczos_step_kernel(int64_t n, double* __restrict__ w, double* __restrict__ v,
                                                            double* __restrict__ vprev, double* __restrict__ alpha,
                                                            double* __restrict__ beta, int j,
                                                            double* __restrict__ state) {
  __shared__ double red[32];
  const double beta_prev = state[0];
  double scale = state[1];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += w[i] * v[i];
  double a = block_sum(acc, red);
  acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double t = w[i] - a * v[i] - beta_prev * vprev[i];
    w[i] = t;
    acc += t * v[i];
  }
  const double a2 = block_sum(acc, red);
  a += a2;
  acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double t = w[i] - a2 * v[i];
    w[i] = t;
    acc += t * t;
  }
  double b = sqrt(block_sum(acc, red));
  scale = fmax(scale, fmax(fabs(a), b));
  const bool ok = b > 1e-13 * scale;
  if (!ok) b = 0.0;
  const double inv = ok ? 1.0 / b : 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    vprev[i] = v[i];
    v[i] = w[i] * inv;
  }
  __syncthreads();  // every thread has read state[] before it is rewritten
  if (threadIdx.x == 0) {
    alpha[j] = a;
    beta[j] = b;
    state[0] = b;
    state[1] = scale;
  }
}

// max(|smallest|, |largest|) eigenvalue of the tridiagonal matrix by multi-section on the Sturm count: warp 0
// brackets the smallest, warp 1 the largest eigenvalue; every round the 32 lanes of a warp probe 32 interior points
// of the bracket (the sequential Sturm recurrence is the latency, so 32 probes cost what one costs) and the
// bracket shrinks 33-fold: 11 rounds instead of 53 bisection steps.
__global__ void __launch_bounds__(64) tridiag_absmax_kernel(const double* __restrict__ alpha,
                                                            const double* __restrict__ beta, int m,
                                                            double* __restrict__ lam_out) {
  __shared__ int cnt[64];
  __shared__ double lo_s[2], hi_s[2];
  const int half = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int target = half == 0 ? 1 : m;  // first x with count(x) >= target
  if (lane == 0) {
    double lo = alpha[0], hi = alpha[0];
    for (int i = 0; i < m; ++i) {  // Gershgorin
      const double r = (i > 0 ? fabs(beta[i - 1]) : 0.0) + (i < m - 1 ? fabs(beta[i]) : 0.0);
      lo = fmin(lo, alpha[i] - r);
      hi = fmax(hi, alpha[i] + r);
    }
    lo_s[half] = lo;
    hi_s[half] = hi;
  }
  __syncthreads();
  for (int round = 0; round < 16; ++round) {
    const double lo = lo_s[half], hi = hi_s[half];
    const double x = lo + (hi - lo) * ((double)(lane + 1) / 33.0);
    cnt[threadIdx.x] = (x > lo && x < hi) ? sturm_count(alpha, beta, m, x) : -1;
    __syncthreads();
    if (lane == 0) {
      double nlo = lo, nhi = hi;
      for (int l = 0; l < 32; ++l) {
        const int c = cnt[half * 32 + l];
        if (c < 0) continue;  // probe not strictly inside the bracket (bracket at the resolution of double)
        const double xl = lo + (hi - lo) * ((double)(l + 1) / 33.0);
        if (c >= target) {
          nhi = xl;
          break;
        }
        nlo = xl;
      }
      lo_s[half] = nlo;
      hi_s[half] = nhi;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    lam_out[0] = fmax(fabs(0.5 * (lo_s[0] + hi_s[0])), fabs(0.5 * (lo_s[1] + hi_s[1])));
}

// ---- shifted systems -----------------------------------------------------------------------------------------
struct ShiftParams {
  double rtol, atol;
  double cs, sn;  // z = c (cs + i sn)
};

__device__ __forceinline__ double cutoff_of(const double* lam, const ShiftParams& p) { return p.rtol * lam[0] + p.atol; }

// M = T - z I (complex128), rhs = b.  A zero cut-off (only possible for T = 0: rtol = atol = 0 is refused) gives
// M = I, rhs = 0, i.e. y = 0 like the reference's where(|lambda| > 0, ., 0); a NaN cut-off propagates.
__global__ void __launch_bounds__(256) shift_build_kernel(const double* __restrict__ T, int64_t n,
                                                          const double* __restrict__ b,
                                                          const double* __restrict__ lam, ShiftParams p,
                                                          cuDoubleComplex* __restrict__ M,
                                                          cuDoubleComplex* __restrict__ rhs) {
  const double c = cutoff_of(lam, p);
  const bool degenerate = (c == 0.0);
  const double zr = c * p.cs, zi = c * p.sn;
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const double* row = T + i * n;
    cuDoubleComplex* out = M + i * n;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      double re = degenerate ? 0.0 : row[j], im = 0.0;
      if (i == j) {
        re = degenerate ? 1.0 : re - zr;
        im = degenerate ? 0.0 : -zi;
      }
      out[j] = make_cuDoubleComplex(re, im);
    }
    if (threadIdx.x == 0) rhs[i] = make_cuDoubleComplex(degenerate ? 0.0 : b[i], 0.0);
  }
}

// x (double-double complex, planar [re.hi | re.lo | im.hi | im.lo]) = first solve
__global__ void __launch_bounds__(256) dd_set_kernel(int64_t n, const cuDoubleComplex* __restrict__ d,
                                                     double* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = d[i].x;
  x[n + i] = 0.0;
  x[2 * n + i] = d[i].y;
  x[3 * n + i] = 0.0;
}

// x += d (correction of one refinement step)
__global__ void __launch_bounds__(256) dd_correct_kernel(int64_t n, const cuDoubleComplex* __restrict__ d,
                                                         double* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const dd re = dd_add_d({x[i], x[n + i]}, d[i].x);
  const dd im = dd_add_d({x[2 * n + i], x[3 * n + i]}, d[i].y);
  x[i] = re.hi;
  x[n + i] = re.lo;
  x[2 * n + i] = im.hi;
  x[3 * n + i] = im.lo;
}

// r = b - (T - z I) x in double-double, rounded to complex128; CTA per row
__global__ void __launch_bounds__(256) dd_residual_kernel(const double* __restrict__ T, int64_t n,
                                                          const double* __restrict__ b,
                                                          const double* __restrict__ lam, ShiftParams p,
                                                          const double* __restrict__ x,
                                                          cuDoubleComplex* __restrict__ r) {
  __shared__ double red[32][4];
  const int64_t i = blockIdx.x;
  const double* row = T + i * n;
  const double *xrh = x, *xrl = x + n, *xih = x + 2 * n, *xil = x + 3 * n;
  dd sr = {0.0, 0.0}, si = {0.0, 0.0};
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double t = row[j];
    sr = dd_fma_acc(sr, t, xrh[j], xrl[j]);
    si = dd_fma_acc(si, t, xih[j], xil[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sr = dd_add(sr, dd_shfl_xor(sr, o));
    si = dd_add(si, dd_shfl_xor(si, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = sr.hi;
    red[threadIdx.x >> 5][1] = sr.lo;
    red[threadIdx.x >> 5][2] = si.hi;
    red[threadIdx.x >> 5][3] = si.lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    dd tr = {0.0, 0.0}, ti = {0.0, 0.0};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      tr = dd_add(tr, {red[w][0], red[w][1]});
      ti = dd_add(ti, {red[w][2], red[w][3]});
    }
    const double c = cutoff_of(lam, p);
    if (c == 0.0) {  // degenerate system M = I, rhs = 0
      r[i] = make_cuDoubleComplex(-xrh[i], -xih[i]);
      return;
    }
    const double zr = c * p.cs, zi = c * p.sn;
    const dd xr = {xrh[i], xrl[i]}, xi = {xih[i], xil[i]};
    // z x = (zr xr - zi xi) + i (zr xi + zi xr)
    const dd zxr = dd_add(dd_mul_d(xr, zr), dd_mul_d(xi, -zi));
    const dd zxi = dd_add(dd_mul_d(xi, zr), dd_mul_d(xr, zi));
    const dd rr = dd_add(dd_add_d(dd_neg(tr), b[i]), zxr);
    const dd ri = dd_add(dd_neg(ti), zxi);
    r[i] = make_cuDoubleComplex(rr.hi + rr.lo, ri.hi + ri.lo);
  }
}

// ydd (+)= Re x   (double-double, planar [hi | lo])
__global__ void __launch_bounds__(256) dd_accum_real_kernel(int64_t n, const double* __restrict__ x,
                                                            double* __restrict__ ydd, int add) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dd v = {x[i], x[n + i]};
  if (add) v = dd_add({ydd[i], ydd[n + i]}, v);
  ydd[i] = v.hi;
  ydd[n + i] = v.lo;
}

__global__ void __launch_bounds__(256) dd_zero_kernel(int64_t n2, double* __restrict__ ydd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) ydd[i] = 0.0;
}

// y = scale * sum_q ydd[q]   (count double-double vectors [count][2][n], summed in order, rounded once)
__global__ void __launch_bounds__(256) dd_sum_scale_kernel(const double* __restrict__ ydd, int count, int64_t n,
                                                           double scale, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dd s = {0.0, 0.0};
  for (int q = 0; q < count; ++q) s = dd_add(s, {ydd[(2 * (int64_t)q) * n + i], ydd[(2 * (int64_t)q + 1) * n + i]});
  y[i] = (s.hi + s.lo) * scale;
}

__global__ void max_info_kernel(int32_t* __restrict__ info, const int32_t* __restrict__ step_info) {
  if (step_info[0] != 0 && info[0] == 0) info[0] = step_info[0];
}
__global__ void zero_info_kernel(int32_t* __restrict__ info) { info[0] = 0; }

// lower triangle of M = T - z I only (zldlt reads nothing else), rhs = b; same degenerate / NaN rules as above
__global__ void __launch_bounds__(256) shift_build_lower_kernel(const double* __restrict__ T, int64_t n,
                                                                const double* __restrict__ b,
                                                                const double* __restrict__ lam, ShiftParams p,
                                                                cuDoubleComplex* __restrict__ M,
                                                                cuDoubleComplex* __restrict__ rhs) {
  const double c = cutoff_of(lam, p);
  const bool degenerate = (c == 0.0);
  const double zr = c * p.cs, zi = c * p.sn;
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const double* row = T + i * n;
    cuDoubleComplex* out = M + i * n;
    for (int64_t j = threadIdx.x; j <= i; j += blockDim.x) {
      double re = degenerate ? 0.0 : row[j], im = 0.0;
      if (i == j) {
        re = degenerate ? 1.0 : re - zr;
        im = degenerate ? 0.0 : -zi;
      }
      out[j] = make_cuDoubleComplex(re, im);
    }
    if (threadIdx.x == 0) rhs[i] = make_cuDoubleComplex(degenerate ? 0.0 : b[i], 0.0);
  }
}

// after the last refinement step: the correction d must be small against the solution x, otherwise the factors are
// useless (the refinement did not contract) and info = code (< 0) says so.  One CTA.
__global__ void __launch_bounds__(256) refine_check_kernel(int64_t n, const cuDoubleComplex* __restrict__ d,
                                                           const double* __restrict__ x, int32_t* __restrict__ info,
                                                           int32_t code) {
  __shared__ double red[2][8];
  double md = 0.0, mx = 0.0;
  bool bad = false;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = fmax(fabs(d[i].x), fabs(d[i].y)), v = fmax(fabs(x[i]), fabs(x[2 * n + i]));
    bad = bad || !(a == a) || !(v == v);
    md = fmax(md, a);
    mx = fmax(mx, v);
  }
  if (bad) md = 1e300;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    md = fmax(md, __shfl_xor_sync(FULL, md, o));
    mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = md;
    red[1][threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      md = fmax(md, red[0][w]);
      mx = fmax(mx, red[1][w]);
    }
    if (md > 1e-2 * mx && info[0] == 0) info[0] = code;
  }
}

struct RationalLayout {
  size_t M, rhs, ipiv, work, x, lanczos, info, total;
  int lwork;
};

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static int rational_layout(int64_t n, RationalLayout* L) {
  cusolverDnHandle_t h;
  int rc = solver_handle(&h);
  if (rc) return rc;
  int lwork = 0;
  cusolverStatus_t s = cusolverDnZgetrf_bufferSize(h, (int)n, (int)n, nullptr, (int)n, &lwork);
  if (s != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolverDnZgetrf_bufferSize failed with status %d", (int)s);
    return QTX_ERR_SOLVER;
  }
  L->lwork = lwork;
  size_t off = 0;
  L->M = off;
  off += align_up((size_t)n * n * sizeof(cuDoubleComplex));
  L->rhs = off;
  off += align_up((size_t)n * sizeof(cuDoubleComplex));
  L->ipiv = off;
  off += align_up((size_t)n * sizeof(int));
  L->work = off;
  off += align_up((size_t)lwork * sizeof(cuDoubleComplex));
  L->x = off;
  off += align_up(4 * (size_t)n * sizeof(double));
  L->lanczos = off;  // [w | v | vprev | alpha | beta | state(2)]
  off += align_up((3 * (size_t)n + 2 * kLanczosMaxSteps + 2) * sizeof(double));
  L->info = off;
  off += 256;
  L->total = off + 256;  // slack for aligning the caller's pointer
  return QTX_OK;
}

}  // namespace qtx

using namespace qtx;

extern "C" size_t qtx_pinv_rational_workspace_size(int64_t n) {
  RationalLayout L;
  if (n <= 0 || n > 46340 || rational_layout(n, &L)) return 0;
  return L.total;
}

extern "C" int qtx_sym_absmax_eig(const double* T, int64_t n, int first_step, int steps, double* lam_out,
                                  void* workspace, size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(T && lam_out && workspace && n > 0 && n <= 46340 && steps > 0 && steps <= kLanczosMaxSteps &&
                  first_step >= 0 && first_step < steps,
              QTX_ERR_INVALID, "qtx_sym_absmax_eig: bad argument");
  RationalLayout L;
  int rc = rational_layout(n, &L);
  if (rc) return rc;
  QTX_REQUIRE(workspace_bytes >= L.total, QTX_ERR_INVALID, "qtx_sym_absmax_eig: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)align_up((size_t)workspace);
  double* w = (double*)(base + L.lanczos);
  double *v = w + n, *vprev = v + n, *alpha = vprev + n, *beta = alpha + kLanczosMaxSteps,
         *state = beta + kLanczosMaxSteps;
  const int m = steps < n ? steps : (int)n;
  if (first_step == 0) {  // otherwise the recurrence continues from the state left in the workspace
    QTX_LAUNCH(lanczos_init_kernel, 1, kLanczosThreads, st, n, v, vprev, state);
    QTX_LAUNCH_CHECK();
  }
  for (int j = first_step; j < m; ++j) {
    rc = qtx_matvec(QTX_F64, T, n, n, n, v, w, stream);
    if (rc) return rc;
    QTX_LAUNCH(lanczos_step_kernel, 1, kLanczosThreads, st, n, w, v, vprev, alpha, beta, j, state);
    QTX_LAUNCH_CHECK();
  }
  QTX_LAUNCH(tridiag_absmax_kernel, 1, 64, st, alpha, beta, m, lam_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_pinv_rational_partial(const double* T, int64_t n, const double* b, double rtol, double atol,
                                         const double* lam, int shift_mask, int refine_steps, double* ydd_inout,
                                         int accumulate, int32_t* info_out, void* workspace, size_t workspace_bytes,
                                         qtx_stream_t stream) {
  QTX_REQUIRE(T && b && lam && ydd_inout && info_out && workspace && n > 0 && n <= 46340 && shift_mask >= 0 &&
                  shift_mask < 8 && refine_steps >= 0 && refine_steps <= 16 && atol >= 0.0,
              QTX_ERR_INVALID, "qtx_pinv_rational_partial: bad argument");
  if (rtol < 0) rtol = 1e-12;  // solver.py:12-21 for float64
  QTX_REQUIRE(rtol > 0.0 || atol > 0.0, QTX_ERR_UNSUPPORTED,
              "qtx_pinv_rational_partial: rtol = atol = 0 is the plain inverse, use qtx_pinv_eig_solve");
  RationalLayout L;
  int rc = rational_layout(n, &L);
  if (rc) return rc;
  QTX_REQUIRE(workspace_bytes >= L.total, QTX_ERR_INVALID, "qtx_pinv_rational_partial: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  cusolverDnHandle_t h;
  rc = solver_handle(&h);
  if (rc) return rc;
  cusolverStatus_t s = cusolverDnSetStream(h, st);
  QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnSetStream failed (%d)", (int)s);
  char* base = (char*)align_up((size_t)workspace);
  cuDoubleComplex* M = (cuDoubleComplex*)(base + L.M);
  cuDoubleComplex* rhs = (cuDoubleComplex*)(base + L.rhs);
  int* ipiv = (int*)(base + L.ipiv);
  cuDoubleComplex* work = (cuDoubleComplex*)(base + L.work);
  double* x = (double*)(base + L.x);
  int32_t* step_info = (int32_t*)(base + L.info);
  const unsigned gn = (unsigned)((n + 255) / 256);
  QTX_LAUNCH(zero_info_kernel, 1, 1, st, info_out);
  QTX_LAUNCH_CHECK();
  if (!accumulate) {
    QTX_LAUNCH(dd_zero_kernel, (unsigned)((2 * n + 255) / 256), 256, st, 2 * n, ydd_inout);
    QTX_LAUNCH_CHECK();
  }
  static const double kCos[3] = {0.86602540378443864676, 0.0, -0.86602540378443864676};  // cos(pi (2k+1)/6)
  static const double kSin[3] = {0.5, 1.0, 0.5};
  for (int k = 0; k < 3; ++k) {
    if (!((shift_mask >> k) & 1)) continue;
    ShiftParams p = {rtol, atol, kCos[k], kSin[k]};
    const unsigned gb = n < 16 * (int64_t)num_sms() ? (unsigned)n : 16u * (unsigned)num_sms();
    QTX_LAUNCH(shift_build_kernel, gb, 256, st, T, n, b, lam, p, M, rhs);
    QTX_LAUNCH_CHECK();
    // the matrix is complex SYMMETRIC, so its row-major image is its column-major image
    s = cusolverDnZgetrf(h, (int)n, (int)n, M, (int)n, work, ipiv, step_info);
    QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnZgetrf failed (%d)", (int)s);
    count_launch();
    QTX_LAUNCH(max_info_kernel, 1, 1, st, info_out, step_info);
    QTX_LAUNCH_CHECK();
    for (int it = 0; it <= refine_steps; ++it) {
      if (it > 0) {
        QTX_LAUNCH(dd_residual_kernel, (unsigned)n, kRowThreads, st, T, n, b, lam, p, x, rhs);
        QTX_LAUNCH_CHECK();
      }
      s = cusolverDnZgetrs(h, CUBLAS_OP_N, (int)n, 1, M, (int)n, ipiv, rhs, (int)n, step_info);
      QTX_REQUIRE(s == CUSOLVER_STATUS_SUCCESS, QTX_ERR_SOLVER, "cusolverDnZgetrs failed (%d)", (int)s);
      count_launch();
      if (it == 0) QTX_LAUNCH(dd_set_kernel, gn, 256, st, n, rhs, x);
      else QTX_LAUNCH(dd_correct_kernel, gn, 256, st, n, rhs, x);
      QTX_LAUNCH_CHECK();
    }
    QTX_LAUNCH(dd_accum_real_kernel, gn, 256, st, n, x, ydd_inout, 1);
    QTX_LAUNCH_CHECK();
  }
  return QTX_OK;
}

// ---- the same partial sums with OWN kernels: complex-symmetric LDL^T (zldlt.cu) instead of cuSOLVER's LU -----------
namespace qtx {

struct LdltLayout {
  size_t M, rhs, x, scratch, info, slot, lanczos, total;  // offsets inside a shift slot; slot = its size
};

static LdltLayout ldlt_layout(int64_t n, int nslots) {
  LdltLayout L;
  size_t off = 0;
  L.M = off;
  off += align_up((size_t)n * n * sizeof(cuDoubleComplex));
  L.rhs = off;
  off += align_up((size_t)n * sizeof(cuDoubleComplex));
  L.x = off;
  off += align_up(4 * (size_t)n * sizeof(double));
  L.scratch = off;
  off += align_up(zldlt_scratch_bytes(n));
  L.info = off;
  off += 256;
  L.slot = off;
  L.lanczos = (size_t)nslots * L.slot;  // [w | v | vprev | alpha | beta | state(2)]: the layout qtx_sym_absmax_eig uses
  L.total = L.lanczos + align_up((3 * (size_t)n + 2 * kLanczosMaxSteps + 2) * sizeof(double)) + 256;
  return L;
}

#ifndef QTX_HOST_EMULATION
// two side streams for running the three shifts of one call concurrently (fork / join with events on the caller's
// stream: the call still only ENQUEUES work that is ordered after, and joined back into, `stream`)
struct ShiftStreams {
  int device = -1;
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
};
static thread_local ShiftStreams g_shift_streams;

static int shift_streams(ShiftStreams** out) {
  int dev = 0;
  QTX_CUDA(cudaGetDevice(&dev));
  ShiftStreams& s = g_shift_streams;
  if (s.device != dev) {
    for (int i = 0; i < 2; ++i) {
      QTX_CUDA(cudaStreamCreateWithFlags(&s.side[i], cudaStreamNonBlocking));
      QTX_CUDA(cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming));
    }
    QTX_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    s.device = dev;
  }
  *out = &s;
  return QTX_OK;
}
#endif

// one shift on one stream: build, factor, solve + refine; leaves the double-double solution in slot.x
static int ldlt_one_shift(const double* T, int64_t n, const double* b, const double* lam, ShiftParams p, int k,
                          int refine_steps, char* slot, const LdltLayout& L, cudaStream_t st) {
  cuDoubleComplex* M = (cuDoubleComplex*)(slot + L.M);
  cuDoubleComplex* rhs = (cuDoubleComplex*)(slot + L.rhs);
  double* x = (double*)(slot + L.x);
  void* scratch = slot + L.scratch;
  int32_t* info = (int32_t*)(slot + L.info);
  const unsigned gn = (unsigned)((n + 255) / 256);
  QTX_LAUNCH(zero_info_kernel, 1, 1, st, info);
  QTX_LAUNCH_CHECK();
  const unsigned gb = n < 16 * (int64_t)num_sms() ? (unsigned)n : 16u * (unsigned)num_sms();
  QTX_LAUNCH(shift_build_lower_kernel, gb, 256, st, T, n, b, lam, p, M, rhs);
  QTX_LAUNCH_CHECK();
  int rc = zldlt_factor(M, n, scratch, info, st);
  if (rc) return rc;
  for (int it = 0; it <= refine_steps; ++it) {
    if (it > 0) {
      QTX_LAUNCH(dd_residual_kernel, (unsigned)n, kRowThreads, st, T, n, b, lam, p, x, rhs);
      QTX_LAUNCH_CHECK();
    }
    rc = zldlt_solve(M, n, rhs, scratch, st);
    if (rc) return rc;
    if (it == 0) QTX_LAUNCH(dd_set_kernel, gn, 256, st, n, rhs, x);
    else QTX_LAUNCH(dd_correct_kernel, gn, 256, st, n, rhs, x);
    QTX_LAUNCH_CHECK();
    if (it == refine_steps && it > 0) {
      QTX_LAUNCH(refine_check_kernel, 1, 256, st, n, rhs, x, info, -(k + 1));
      QTX_LAUNCH_CHECK();
    }
  }
  return QTX_OK;
}

}  // namespace qtx

extern "C" size_t qtx_pinv_ldlt_workspace_size(int64_t n, int nshifts) {
  if (n <= 0 || n > 46340 || nshifts < 1 || nshifts > 3) return 0;
  return ldlt_layout(n, nshifts).total;
}

// Lanczos for max|lambda| in the workspace of qtx_pinv_ldlt_workspace_size(n, nshifts) (same recurrence, same
// continuation protocol as qtx_sym_absmax_eig)
extern "C" int qtx_sym_absmax_eig_ws(const double* T, int64_t n, int first_step, int steps, double* lam_out,
                                     void* workspace, size_t workspace_bytes, int nshifts, qtx_stream_t stream) {
  QTX_REQUIRE(T && lam_out && workspace && n > 0 && n <= 46340 && steps > 0 && steps <= kLanczosMaxSteps &&
                  first_step >= 0 && first_step < steps && nshifts >= 1 && nshifts <= 3,
              QTX_ERR_INVALID, "qtx_sym_absmax_eig_ws: bad argument");
  const LdltLayout L = ldlt_layout(n, nshifts);
  QTX_REQUIRE(workspace_bytes >= L.total, QTX_ERR_INVALID, "qtx_sym_absmax_eig_ws: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)align_up((size_t)workspace);
  double* w = (double*)(base + L.lanczos);
  double *v = w + n, *vprev = v + n, *alpha = vprev + n, *beta = alpha + kLanczosMaxSteps,
         *state = beta + kLanczosMaxSteps;
  const int m = steps < n ? steps : (int)n;
  if (first_step == 0) {
    QTX_LAUNCH(lanczos_init_kernel, 1, kLanczosThreads, st, n, v, vprev, state);
    QTX_LAUNCH_CHECK();
  }
  for (int j = first_step; j < m; ++j) {
    int rc = qtx_matvec(QTX_F64, T, n, n, n, v, w, stream);
    if (rc) return rc;
    QTX_LAUNCH(lanczos_step_kernel, 1, kLanczosThreads, st, n, w, v, vprev, alpha, beta, j, state);
    QTX_LAUNCH_CHECK();
  }
  QTX_LAUNCH(tridiag_absmax_kernel, 1, 64, st, alpha, beta, m, lam_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_pinv_ldlt_partial(const double* T, int64_t n, const double* b, double rtol, double atol,
                                     const double* lam, int shift_mask, int refine_steps, double* ydd_inout,
                                     int accumulate, int32_t* info_out, void* workspace, size_t workspace_bytes,
                                     qtx_stream_t stream) {
  QTX_REQUIRE(T && b && lam && ydd_inout && info_out && workspace && n > 0 && n <= 46340 && shift_mask >= 0 &&
                  shift_mask < 8 && refine_steps >= 0 && refine_steps <= 16 && atol >= 0.0,
              QTX_ERR_INVALID, "qtx_pinv_ldlt_partial: bad argument");
  if (rtol < 0) rtol = 1e-12;  // solver.py:12-21 for float64
  QTX_REQUIRE(rtol > 0.0 || atol > 0.0, QTX_ERR_UNSUPPORTED,
              "qtx_pinv_ldlt_partial: rtol = atol = 0 is the plain inverse, use qtx_pinv_eig_solve");
  int nshifts = 0;
  for (int k = 0; k < 3; ++k) nshifts += (shift_mask >> k) & 1;
  const LdltLayout L = ldlt_layout(n, nshifts > 0 ? nshifts : 1);
  QTX_REQUIRE(workspace_bytes >= L.total, QTX_ERR_INVALID, "qtx_pinv_ldlt_partial: workspace too small for %d shifts",
              nshifts);
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)align_up((size_t)workspace);
  const unsigned gn = (unsigned)((n + 255) / 256);
  QTX_LAUNCH(zero_info_kernel, 1, 1, st, info_out);
  QTX_LAUNCH_CHECK();
  if (!accumulate) {
    QTX_LAUNCH(dd_zero_kernel, (unsigned)((2 * n + 255) / 256), 256, st, 2 * n, ydd_inout);
    QTX_LAUNCH_CHECK();
  }
  if (nshifts == 0) return QTX_OK;
  static const double kCos[3] = {0.86602540378443864676, 0.0, -0.86602540378443864676};  // cos(pi (2k+1)/6)
  static const double kSin[3] = {0.5, 1.0, 0.5};
#ifndef QTX_HOST_EMULATION
  ShiftStreams* ss = nullptr;
  if (nshifts > 1) {
    int rc = shift_streams(&ss);
    if (rc) return rc;
    QTX_CUDA(cudaEventRecord(ss->fork, st));
  }
#endif
  int slot = 0;
  for (int k = 0; k < 3; ++k) {
    if (!((shift_mask >> k) & 1)) continue;
    const ShiftParams p = {rtol, atol, kCos[k], kSin[k]};
    cudaStream_t sk = st;
#ifndef QTX_HOST_EMULATION
    if (slot > 0) {
      sk = ss->side[slot - 1];
      QTX_CUDA(cudaStreamWaitEvent(sk, ss->fork, 0));
    }
#endif
    int rc = ldlt_one_shift(T, n, b, lam, p, k, refine_steps, base + (size_t)slot * L.slot, L, sk);
    if (rc) return rc;
#ifndef QTX_HOST_EMULATION
    if (slot > 0) {
      QTX_CUDA(cudaEventRecord(ss->join[slot - 1], sk));
      QTX_CUDA(cudaStreamWaitEvent(st, ss->join[slot - 1], 0));
    }
#endif
    ++slot;
  }
  for (int q = 0; q < nshifts; ++q) {  // the shifts are summed in a fixed order on the caller's stream
    char* sl = base + (size_t)q * L.slot;
    QTX_LAUNCH(dd_accum_real_kernel, gn, 256, st, n, (const double*)(sl + L.x), ydd_inout, 1);
    QTX_LAUNCH_CHECK();
    QTX_LAUNCH(max_info_kernel, 1, 1, st, info_out, (const int32_t*)(sl + L.info));
    QTX_LAUNCH_CHECK();
  }
  return QTX_OK;
}

extern "C" int qtx_dd_sum_scale(const double* ydd, int count, int64_t n, double scale, double* y_out,
                                qtx_stream_t stream) {
  QTX_REQUIRE(ydd && y_out && count > 0 && n > 0, QTX_ERR_INVALID, "qtx_dd_sum_scale: bad argument");
  QTX_LAUNCH(dd_sum_scale_kernel, (unsigned)((n + 255) / 256), 256, (cudaStream_t)stream, ydd, count, n, scale, y_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
