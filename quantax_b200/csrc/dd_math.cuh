// Double-double arithmetic (error-free transformations) and the Sturm-count bisection used by pinv_rational.cu.
// Host + device: tests/test_dd_math_cpu.py compiles this header with g++ and checks it against exact rational
// arithmetic, so the arithmetic of the kernels is verified on the CPU from the same source.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define QTX_HD __host__ __device__ __forceinline__
#else
#define QTX_HD inline
#endif

namespace qtx {

// rounded operations that the compiler must neither contract into FMAs nor reassociate
#if defined(__CUDA_ARCH__)
#define QTX_ADD(a, b) __dadd_rn((a), (b))
#define QTX_SUB(a, b) __dsub_rn((a), (b))
#define QTX_MUL(a, b) __dmul_rn((a), (b))
#define QTX_FMA(a, b, c) __fma_rn((a), (b), (c))
#else  // host: compile with -ffp-contract=off (no -ffast-math)
#define QTX_ADD(a, b) ((a) + (b))
#define QTX_SUB(a, b) ((a) - (b))
#define QTX_MUL(a, b) ((a) * (b))
#define QTX_FMA(a, b, c) fma((a), (b), (c))
#endif

struct dd {
  double hi, lo;
};

QTX_HD dd two_sum(double a, double b) {
  const double s = QTX_ADD(a, b);
  const double bb = QTX_SUB(s, a);
  const double e = QTX_ADD(QTX_SUB(a, QTX_SUB(s, bb)), QTX_SUB(b, bb));
  return {s, e};
}
QTX_HD dd quick_two_sum(double a, double b) {  // |a| >= |b|
  const double s = QTX_ADD(a, b);
  return {s, QTX_SUB(b, QTX_SUB(s, a))};
}
QTX_HD dd two_prod(double a, double b) {
  const double p = QTX_MUL(a, b);
  return {p, QTX_FMA(a, b, -p)};
}
QTX_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s.lo = QTX_ADD(s.lo, t.hi);
  s = quick_two_sum(s.hi, s.lo);
  s.lo = QTX_ADD(s.lo, t.lo);
  return quick_two_sum(s.hi, s.lo);
}
QTX_HD dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  s.lo = QTX_ADD(s.lo, a.lo);
  return quick_two_sum(s.hi, s.lo);
}
QTX_HD dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo = QTX_FMA(a.lo, b, p.lo);
  return quick_two_sum(p.hi, p.lo);
}
QTX_HD dd dd_neg(dd a) { return {-a.hi, -a.lo}; }

// t * (x_hi + x_lo) added to the double-double accumulator s (one term of the residual row sums)
QTX_HD dd dd_fma_acc(dd s, double t, double xh, double xl) {
  dd p = two_prod(t, xh);
  p.lo = QTX_FMA(t, xl, p.lo);
  return dd_add(s, p);
}

// number of eigenvalues below x of the symmetric tridiagonal matrix (alpha [m], beta [m-1])
QTX_HD int sturm_count(const double* alpha, const double* beta, int m, double x) {
  int cnt = 0;
  double d = 1.0;
  for (int i = 0; i < m; ++i) {
    const double off = i > 0 ? beta[i - 1] * beta[i - 1] : 0.0;
    d = (alpha[i] - x) - off / d;
    if (d == 0.0) d = 1e-300;
    if (d < 0.0) ++cnt;
  }
  return cnt;
}

// the target-th smallest eigenvalue (1-based) by bisection between the Gershgorin bounds
QTX_HD double tridiag_eigenvalue(const double* alpha, const double* beta, int m, int target) {
  double lo = alpha[0], hi = alpha[0];
  for (int i = 0; i < m; ++i) {
    const double r = (i > 0 ? fabs(beta[i - 1]) : 0.0) + (i < m - 1 ? fabs(beta[i]) : 0.0);
    lo = fmin(lo, alpha[i] - r);
    hi = fmax(hi, alpha[i] + r);
  }
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (sturm_count(alpha, beta, m, mid) >= target) hi = mid;
    else lo = mid;
  }
  return 0.5 * (lo + hi);
}

}  // namespace qtx
