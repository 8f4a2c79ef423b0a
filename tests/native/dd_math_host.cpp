// Host build of quantax_b200/csrc/dd_math.cuh for tests/test_dd_math_cpu.py (g++ -O2 -ffp-contract=off):
// plain C wrappers around the double-double primitives and the Sturm bisection used by csrc/pinv_rational.cu.
#include "dd_math.cuh"

using namespace qtx;

extern "C" {
void h_two_sum(double a, double b, double* out) { dd r = two_sum(a, b); out[0] = r.hi; out[1] = r.lo; }
void h_two_prod(double a, double b, double* out) { dd r = two_prod(a, b); out[0] = r.hi; out[1] = r.lo; }
void h_dd_add(double ah, double al, double bh, double bl, double* out) {
  dd r = dd_add({ah, al}, {bh, bl}); out[0] = r.hi; out[1] = r.lo;
}
void h_dd_add_d(double ah, double al, double b, double* out) { dd r = dd_add_d({ah, al}, b); out[0] = r.hi; out[1] = r.lo; }
void h_dd_mul_d(double ah, double al, double b, double* out) { dd r = dd_mul_d({ah, al}, b); out[0] = r.hi; out[1] = r.lo; }
// sum_j t[j] * (xh[j] + xl[j]) accumulated like one thread of dd_residual_kernel
void h_dd_dot(const double* t, const double* xh, const double* xl, int n, double* out) {
  dd s = {0.0, 0.0};
  for (int j = 0; j < n; ++j) s = dd_fma_acc(s, t[j], xh[j], xl[j]);
  out[0] = s.hi; out[1] = s.lo;
}
int h_sturm_count(const double* alpha, const double* beta, int m, double x) { return sturm_count(alpha, beta, m, x); }
double h_tridiag_eigenvalue(const double* alpha, const double* beta, int m, int target) {
  return tridiag_eigenvalue(alpha, beta, m, target);
}
}
