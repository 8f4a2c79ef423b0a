// jax.ffi (XLA FFI) custom-call handlers over the C ABI of libqtx_b200 (include/qtx_b200.h): the thin layer
// BASELINE.json's north_star names.  One handler per hot-path entry point, each a 1:1 translation
//   ffi::Buffer<T>            -> device pointer + dimensions
//   ffi::PlatformStream<...>  -> the stream XLA runs the call on
//   attributes                -> scalar arguments
//   ffi::Error                <- qtx status code + qtx_last_error()
// Scratch memory is an extra RESULT buffer of the call (jax.ffi.ffi_call allocates it; size from the matching
// qtx_*_workspace_size query), so no handler allocates.  The handlers only enqueue work.
//
// Reference call sites these handlers replace (the bodies of the jitted functions): sampler/metropolis.py:246-322
// (sweep), operator/operator.py:510-562 (Oloc), state/variational.py:424-511 (jacobian), optimizer/sr.py:74-113 and
// optimizer/solver.py:128-149 (Obar, solve), state/variational.py:558-579 (update).
//
// Build: make -C ffi [XLA_FFI_INCLUDE=<jaxlib include dir>].  Without a jaxlib the file is compiled against
// ffi/stub/xla/ffi/api/ffi.h, a restatement of the public API that only type-checks the bindings.
#include <cuda_runtime_api.h>

#include <string>

#include "qtx_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  std::string msg = std::string(what) + " failed (" + std::to_string(rc) + "): " + qtx_last_error();
  return rc == QTX_ERR_INVALID ? ffi::Error::InvalidArgument(msg) : ffi::Error::Internal(msg);
}

template <typename B>
int64_t Dim(const B& b, size_t i) {
  return b.dimensions()[i];
}

// ---- RBM_Dense: whole Metropolis sweep (sampler/metropolis.py:246-322 with shallow_nets.py:87-108) ---------------------
// W [M, N], b [M] float32; spins_in [ns, N]; nbr [N, max_nb] (exchange) -> spins_out, logabs, logabs_chain, naccept
ffi::Error RbmSweepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> W, ffi::Buffer<ffi::F32> b,
                        ffi::Buffer<ffi::S8> spins_in, ffi::Buffer<ffi::S32> nbr, ffi::ResultBuffer<ffi::S8> spins_out,
                        ffi::ResultBuffer<ffi::F64> logabs, ffi::ResultBuffer<ffi::F64> logabs_chain,
                        ffi::ResultBuffer<ffi::S32> naccept, ffi::ResultBuffer<ffi::U8> workspace, int64_t nsweeps,
                        int64_t kind, int64_t hop, double reweight, int64_t seed, int64_t step0, int64_t chain0) {
  const int M = (int)Dim(W, 0), N = (int)Dim(W, 1);
  const int64_t ns = Dim(spins_in, 0);
  cudaError_t e = cudaMemcpyAsync(spins_out->typed_data(), spins_in.typed_data(), spins_in.size_bytes(),
                                  cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
  const int max_nb = kind == QTX_SPIN_EXCHANGE ? (int)Dim(nbr, 1) : 0;
  return Status(qtx_rbm_sweep(QTX_F32, W.typed_data(), b.typed_data(), N, M, spins_out->typed_data(), ns, (int)nsweeps,
                              (int)kind, nbr.typed_data(), max_nb, (int)hop, reweight, nullptr, nullptr, nullptr,
                              (uint64_t)seed, (uint64_t)step0, (uint64_t)chain0, logabs->typed_data(),
                              logabs_chain->typed_data(), naccept->typed_data(), nullptr, workspace->untyped_data(),
                              workspace->size_bytes(), stream),
                "qtx_rbm_sweep");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_rbm_sweep, RbmSweepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("nsweeps")
                                  .Attr<int64_t>("kind")
                                  .Attr<int64_t>("hop")
                                  .Attr<double>("reweight")
                                  .Attr<int64_t>("seed")
                                  .Attr<int64_t>("step0")
                                  .Attr<int64_t>("chain0"));

// ---- RBM_Dense: fused local energies (operator/operator.py:510-562) -----------------------------------------------------
ffi::Error RbmOlocImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> W, ffi::Buffer<ffi::F32> b, ffi::Buffer<ffi::S8> spins,
                       ffi::Buffer<ffi::F64> term_coef, ffi::Buffer<ffi::U16> term_sites, ffi::Buffer<ffi::U8> term_ops,
                       ffi::ResultBuffer<ffi::F64> eloc, ffi::ResultBuffer<ffi::S32> nconn,
                       ffi::ResultBuffer<ffi::U8> workspace) {
  const int M = (int)Dim(W, 0), N = (int)Dim(W, 1);
  return Status(qtx_rbm_oloc(QTX_F32, W.typed_data(), b.typed_data(), N, M, spins.typed_data(), Dim(spins, 0),
                             term_coef.typed_data(), term_sites.typed_data(), term_ops.typed_data(),
                             (int)Dim(term_coef, 0), eloc->typed_data(), nconn->typed_data(), workspace->untyped_data(),
                             workspace->size_bytes(), stream),
                "qtx_rbm_oloc");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_rbm_oloc, RbmOlocImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U16>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// ---- RBM_Dense: Jacobian written centred and scaled (state/variational.py:424-511 + optimizer/sr.py:74-88) ---------------
ffi::Error RbmJacobianImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> W, ffi::Buffer<ffi::F32> b,
                           ffi::Buffer<ffi::S8> spins, ffi::Buffer<ffi::F64> col_mean, ffi::Buffer<ffi::F64> row_scale,
                           ffi::ResultBuffer<ffi::F64> obar) {
  const int M = (int)Dim(W, 0), N = (int)Dim(W, 1);
  return Status(qtx_rbm_jacobian(QTX_F32, W.typed_data(), b.typed_data(), N, M, spins.typed_data(), Dim(spins, 0), QTX_F64,
                                 obar->typed_data(), Dim(*obar, 1), col_mean.typed_data(), row_scale.typed_data(), nullptr,
                                 stream),
                "qtx_rbm_jacobian");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_rbm_jacobian, RbmJacobianImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// ---- ResConv: batched forward on the tcgen05 tower (model/conv_nets.py:78-183) --------------------------------------------
// params: flat float32 vector in ravel_pytree order; spins [ns, lx * ly]
ffi::Error ResconvForwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::S8> spins,
                              ffi::ResultBuffer<ffi::F64> significand, ffi::ResultBuffer<ffi::F64> exponent,
                              ffi::ResultBuffer<ffi::U8> workspace, int64_t nblocks, int64_t channels, int64_t lx,
                              int64_t ly, int64_t kh, int64_t kw, int64_t final_act) {
  return Status(qtx_resconv_forward(QTX_F32, params.typed_data(), (int)nblocks, (int)channels, (int)lx, (int)ly, (int)kh,
                                    (int)kw, (int)final_act, spins.typed_data(), Dim(spins, 0), significand->typed_data(),
                                    exponent->typed_data(), workspace->untyped_data(), workspace->size_bytes(), stream),
                "qtx_resconv_forward");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_resconv_forward, ResconvForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("nblocks")
                                  .Attr<int64_t>("channels")
                                  .Attr<int64_t>("lx")
                                  .Attr<int64_t>("ly")
                                  .Attr<int64_t>("kh")
                                  .Attr<int64_t>("kw")
                                  .Attr<int64_t>("final_act"));

// ---- ResConv: whole Metropolis sweep, device resident (sampler/metropolis.py:246-275, full forward per proposal) ------------
ffi::Error ResconvSweepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::S8> spins_in,
                            ffi::Buffer<ffi::S32> nbr, ffi::ResultBuffer<ffi::S8> spins_out,
                            ffi::ResultBuffer<ffi::F64> significand, ffi::ResultBuffer<ffi::F64> exponent,
                            ffi::ResultBuffer<ffi::S32> naccept, ffi::ResultBuffer<ffi::U8> workspace, int64_t nblocks,
                            int64_t channels, int64_t lx, int64_t ly, int64_t kh, int64_t kw, int64_t final_act,
                            int64_t nsweeps, int64_t kind, int64_t hop, double reweight, int64_t seed, int64_t step0,
                            int64_t chain0) {
  cudaError_t e = cudaMemcpyAsync(spins_out->typed_data(), spins_in.typed_data(), spins_in.size_bytes(),
                                  cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
  const int max_nb = kind == QTX_SPIN_EXCHANGE ? (int)Dim(nbr, 1) : 0;
  return Status(qtx_resconv_sweep(QTX_F32, params.typed_data(), (int)nblocks, (int)channels, (int)lx, (int)ly, (int)kh,
                                  (int)kw, (int)final_act, spins_out->typed_data(), Dim(spins_in, 0), (int)nsweeps,
                                  (int)kind, nbr.typed_data(), max_nb, (int)hop, reweight, (uint64_t)seed, (uint64_t)step0,
                                  (uint64_t)chain0, significand->typed_data(), exponent->typed_data(),
                                  naccept->typed_data(), workspace->untyped_data(), workspace->size_bytes(), stream),
                "qtx_resconv_sweep");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_resconv_sweep, ResconvSweepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("nblocks")
                                  .Attr<int64_t>("channels")
                                  .Attr<int64_t>("lx")
                                  .Attr<int64_t>("ly")
                                  .Attr<int64_t>("kh")
                                  .Attr<int64_t>("kw")
                                  .Attr<int64_t>("final_act")
                                  .Attr<int64_t>("nsweeps")
                                  .Attr<int64_t>("kind")
                                  .Attr<int64_t>("hop")
                                  .Attr<double>("reweight")
                                  .Attr<int64_t>("seed")
                                  .Attr<int64_t>("step0")
                                  .Attr<int64_t>("chain0"));

// ---- ResConv: per-sample log-derivatives (state/variational.py:424-511) ---------------------------------------------------
ffi::Error ResconvJacobianImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::S8> spins,
                               ffi::ResultBuffer<ffi::F64> jac, ffi::ResultBuffer<ffi::U8> workspace, int64_t nblocks,
                               int64_t channels, int64_t lx, int64_t ly, int64_t kh, int64_t kw, int64_t final_act) {
  return Status(qtx_resconv_jacobian(QTX_F32, params.typed_data(), (int)nblocks, (int)channels, (int)lx, (int)ly, (int)kh,
                                     (int)kw, (int)final_act, spins.typed_data(), Dim(spins, 0), QTX_F64,
                                     jac->typed_data(), Dim(*jac, 1), nullptr, nullptr, workspace->untyped_data(),
                                     workspace->size_bytes(), stream),
                "qtx_resconv_jacobian");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_resconv_jacobian, ResconvJacobianImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("nblocks")
                                  .Attr<int64_t>("channels")
                                  .Attr<int64_t>("lx")
                                  .Attr<int64_t>("ly")
                                  .Attr<int64_t>("kh")
                                  .Attr<int64_t>("kw")
                                  .Attr<int64_t>("final_act"));

// ---- generic Oloc reduction: Eloc[s] += sum_c H_c psi(s'_c) / psi(s) (operator/operator.py:168-184) -------------------------
ffi::Error OlocReduceImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> segment, ffi::Buffer<ffi::F64> H,
                          ffi::Buffer<ffi::F64> mult_conn, ffi::Buffer<ffi::F64> expo_conn, ffi::Buffer<ffi::F64> mult,
                          ffi::Buffer<ffi::F64> expo, ffi::Buffer<ffi::F64> diag, ffi::ResultBuffer<ffi::F64> eloc) {
  cudaError_t e = cudaMemcpyAsync(eloc->typed_data(), diag.typed_data(), diag.size_bytes(), cudaMemcpyDeviceToDevice,
                                  stream);
  if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
  return Status(qtx_oloc_reduce(segment.typed_data(), H.typed_data(), mult_conn.typed_data(), expo_conn.typed_data(),
                                Dim(segment, 0), mult.typed_data(), expo.typed_data(), Dim(mult, 0), eloc->typed_data(),
                                stream),
                "qtx_oloc_reduce");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_oloc_reduce, OlocReduceImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// ---- Ebar, energy, VarE (optimizer/sr.py:180-195) ----------------------------------------------------------------------------
ffi::Error EbarImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> eloc, ffi::Buffer<ffi::F64> rw,
                    ffi::ResultBuffer<ffi::F64> ebar, ffi::ResultBuffer<ffi::F64> stats) {
  return Status(qtx_ebar(eloc.typed_data(), rw.typed_data(), Dim(eloc, 0), ebar->typed_data(), stats->typed_data(), stream),
                "qtx_ebar");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_ebar, EbarImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// ---- Obar = (O - mean) sqrt(rw / Ns) (optimizer/sr.py:74-88): column mean, then centre + scale a copy -------------------------
ffi::Error CenterScaleImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> omat, ffi::Buffer<ffi::F64> scale,
                           ffi::ResultBuffer<ffi::F64> obar, ffi::ResultBuffer<ffi::F64> mean) {
  const int64_t ns = Dim(omat, 0), np = Dim(omat, 1);
  int rc = qtx_colmean(QTX_F64, omat.typed_data(), ns, np, np, nullptr, mean->typed_data(), stream);
  if (rc) return Status(rc, "qtx_colmean");
  cudaError_t e = cudaMemcpyAsync(obar->typed_data(), omat.typed_data(), omat.size_bytes(), cudaMemcpyDeviceToDevice,
                                  stream);
  if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
  return Status(qtx_center_scale(QTX_F64, obar->typed_data(), ns, np, np, mean->typed_data(), scale.typed_data(), stream),
                "qtx_center_scale");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_center_scale, CenterScaleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// ---- T = Obar Obar^T on tcgen05 (optimizer/solver.py:139) -----------------------------------------------------------------------
ffi::Error GramImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> obar, ffi::ResultBuffer<ffi::F64> T,
                    ffi::ResultBuffer<ffi::U8> workspace, int64_t nslices) {
  const int64_t ns = Dim(obar, 0), np = Dim(obar, 1);
  return Status(qtx_gram(QTX_F64, obar.typed_data(), ns, np, np, (int)nslices, T->typed_data(), 0, workspace->untyped_data(),
                         workspace->size_bytes(), stream),
                "qtx_gram");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_gram, GramImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("nslices"));

// ---- y = f(T) b, the soft pseudo-inverse of optimizer/solver.py:94-111,142-146 without eigh ------------------------------------
// own LDL^T kernels; max|lambda| from `lanczos_steps` Lanczos steps (no host read-back inside a custom call)
ffi::Error PinvSolveImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> T, ffi::Buffer<ffi::F64> b,
                         ffi::ResultBuffer<ffi::F64> y, ffi::ResultBuffer<ffi::S32> info,
                         ffi::ResultBuffer<ffi::F64> scratch /* [2 n + 1]: ydd, lambda */,
                         ffi::ResultBuffer<ffi::U8> workspace, double rtol, double atol, int64_t lanczos_steps,
                         int64_t refine_steps) {
  const int64_t n = Dim(T, 0);
  double* ydd = scratch->typed_data();
  double* lam = ydd + 2 * n;
  int rc = qtx_sym_absmax_eig_ws(T.typed_data(), n, 0, (int)lanczos_steps, lam, workspace->untyped_data(),
                                 workspace->size_bytes(), 3, stream);
  if (rc) return Status(rc, "qtx_sym_absmax_eig_ws");
  rc = qtx_pinv_ldlt_partial(T.typed_data(), n, b.typed_data(), rtol, atol, lam, 7, (int)refine_steps, ydd, 0,
                             info->typed_data(), workspace->untyped_data(), workspace->size_bytes(), stream);
  if (rc) return Status(rc, "qtx_pinv_ldlt_partial");
  return Status(qtx_dd_sum_scale(ydd, 1, n, 1.0 / 3.0, y->typed_data(), stream), "qtx_dd_sum_scale");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_pinv_solve, PinvSolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<int64_t>("lanczos_steps")
                                  .Attr<int64_t>("refine_steps"));

// ---- x = Obar^T y (optimizer/solver.py:146) ------------------------------------------------------------------------------------
ffi::Error MatvecTImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> obar, ffi::Buffer<ffi::F64> y,
                       ffi::ResultBuffer<ffi::F64> x) {
  const int64_t ns = Dim(obar, 0), np = Dim(obar, 1);
  return Status(qtx_matvec_t(QTX_F64, obar.typed_data(), ns, np, np, y.typed_data(), x->typed_data(), 0, stream),
                "qtx_matvec_t");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_matvec_t, MatvecTImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

// ---- distributed MinSR solve over the library's own communicator (optimizer/solver.py:128-149 under GSPMD) -----------------------
// `comm` is the qtx_comm_t the host created once (qtx_comm_init / qtx_comm_adopt), passed as an integer attribute;
// lanczos_steps < 0 runs exactly that many Lanczos steps without a host read-back inside the custom call, > 0 is the
// adaptive run of the single-GPU solve (include/qtx_b200.h)
ffi::Error MinsrSolveDistImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> obar_local, ffi::Buffer<ffi::F64> ebar_local,
                              ffi::ResultBuffer<ffi::F64> x, ffi::ResultBuffer<ffi::S32> info,
                              ffi::ResultBuffer<ffi::U8> workspace, int64_t comm, double rtol, double atol,
                              int64_t nslices, int64_t lanczos_steps, int64_t refine_steps) {
  const int64_t nl = Dim(obar_local, 0), np = Dim(obar_local, 1);
  return Status(qtx_minsr_solve_dist(reinterpret_cast<qtx_comm_t>(comm), QTX_F64, obar_local.typed_data(), nl, np, np,
                                     ebar_local.typed_data(), rtol, atol, (int)nslices, (int)lanczos_steps,
                                     (int)refine_steps, x->typed_data(), info->typed_data(), workspace->untyped_data(),
                                     workspace->size_bytes(), stream),
                "qtx_minsr_solve_dist");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_minsr_solve_dist, MinsrSolveDistImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("comm")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<int64_t>("nslices")
                                  .Attr<int64_t>("lanczos_steps")
                                  .Attr<int64_t>("refine_steps"));

// ---- params <- params - lr * step, skipped when the step is not finite (state/variational.py:558-579) ------------------------------
ffi::Error ApplyUpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F64> step,
                           ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::S32> applied, double lr) {
  cudaError_t e = cudaMemcpyAsync(params_out->typed_data(), params.typed_data(), params.size_bytes(),
                                  cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
  return Status(qtx_apply_update(QTX_F32, params_out->typed_data(), step.typed_data(), lr, Dim(step, 0),
                                 applied->typed_data(), stream),
                "qtx_apply_update");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(qtx_ffi_apply_update, ApplyUpdateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<double>("lr"));

}  // namespace
