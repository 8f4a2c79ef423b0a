// Compiles quantax_b200/csrc/pinv_rational.cu for the CPU (tests/native/cuda_emu.h: one std::thread per CUDA thread;
// cuda_emu_host.h: launch macro, error macros, a LAPACK-style stand-in for cuSOLVER), so that the ENTRY POINTS
// qtx_sym_absmax_eig / qtx_pinv_rational_partial / qtx_dd_sum_scale run as they are; the emu_* wrappers below run
// single kernels with a chosen launch geometry.  TEST INFRASTRUCTURE ONLY.
#define QTX_HOST_EMULATION 1
#include "zldlt.cu"  // the own-kernel route of pinv_rational.cu (qtx_pinv_ldlt_partial) links against it
#include "pinv_rational.cu"

// what the entry points of pinv_rational.cu link against inside the library
namespace qtx {
int solver_handle(cusolverDnHandle_t* h) {
  static EmuSolver solver;
  *h = &solver;
  return QTX_OK;
}
}  // namespace qtx
extern "C" int qtx_matvec(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, const double* x, double* v_out,
                          qtx_stream_t) {
  if (dtype != QTX_F64 || !A || !x || !v_out) return QTX_ERR_INVALID;
  const double* a = (const double*)A;
  for (int64_t r = 0; r < ns; ++r) {
    double acc = 0.0;
    for (int64_t k = 0; k < np; ++k) acc += a[r * ld + k] * x[k];
    v_out[r] = acc;
  }
  return QTX_OK;
}
extern "C" const char* emu_last_error() { return qtx::g_emu_error; }

extern "C" {

// steps [first_step, steps) of the Lanczos recurrence (state = [w | v | vprev | alpha | beta | state(2)] as in the
// workspace layout of the library), then the bisection: the body of qtx_sym_absmax_eig with a host matvec
void emu_sym_absmax_eig(const double* T, int64_t n, int first_step, int steps, unsigned block, double* work,
                        double* lam_out) {  // the library launches 1024 threads; any multiple of 32 must do
  double* w = work;
  double *v = w + n, *vprev = v + n, *alpha = vprev + n, *beta = alpha + kLanczosMaxSteps,
         *state = beta + kLanczosMaxSteps;
  const int m = steps < n ? steps : (int)n;
  if (first_step == 0) emu_launch(1, block, [&] { lanczos_init_kernel(n, v, vprev, state); });
  for (int j = first_step; j < m; ++j) {
    for (int64_t r = 0; r < n; ++r) {  // qtx_matvec
      double acc = 0.0;
      for (int64_t k = 0; k < n; ++k) acc += T[r * n + k] * v[k];
      w[r] = acc;
    }
    emu_launch(1, block, [&] { lanczos_step_kernel(n, w, v, vprev, alpha, beta, j, state); });
  }
  emu_launch(1, 64, [&] { tridiag_absmax_kernel(alpha, beta, m, lam_out); });
}

void emu_shift_build(const double* T, int64_t n, const double* b, const double* lam, double rtol, double atol, double cs,
                     double sn, unsigned grid, double* M_c128, double* rhs_c128) {
  ShiftParams p = {rtol, atol, cs, sn};
  emu_launch(grid, 256, [&] { shift_build_kernel(T, n, b, lam, p, (cuDoubleComplex*)M_c128, (cuDoubleComplex*)rhs_c128); });
}

void emu_dd_set(int64_t n, const double* d_c128, double* x) {
  emu_launch((unsigned)((n + 255) / 256), 256, [&] { dd_set_kernel(n, (const cuDoubleComplex*)d_c128, x); });
}

void emu_dd_correct(int64_t n, const double* d_c128, double* x) {
  emu_launch((unsigned)((n + 255) / 256), 256, [&] { dd_correct_kernel(n, (const cuDoubleComplex*)d_c128, x); });
}

void emu_dd_residual(const double* T, int64_t n, const double* b, const double* lam, double rtol, double atol, double cs,
                     double sn, const double* x, double* r_c128) {
  ShiftParams p = {rtol, atol, cs, sn};
  emu_launch((unsigned)n, 256, [&] { dd_residual_kernel(T, n, b, lam, p, x, (cuDoubleComplex*)r_c128); });  // library size
}

void emu_dd_zero(int64_t n2, double* ydd) {
  emu_launch((unsigned)((n2 + 255) / 256), 256, [&] { dd_zero_kernel(n2, ydd); });
}

void emu_dd_accum_real(int64_t n, const double* x, double* ydd) {
  emu_launch((unsigned)((n + 255) / 256), 256, [&] { dd_accum_real_kernel(n, x, ydd, 1); });
}

void emu_dd_sum_scale(const double* ydd, int count, int64_t n, double scale, double* y) {
  emu_launch((unsigned)((n + 255) / 256), 256, [&] { dd_sum_scale_kernel(ydd, count, n, scale, y); });
}

int emu_lanczos_max_steps() { return kLanczosMaxSteps; }
}
