// Whole Metropolis sweep of a ResConv state behind ONE C-ABI call (quantax/sampler/metropolis.py:246-275 for a model
// without local updates: propose -> full forward of the proposed chains -> accept, `nsweeps` times).  The host makes one
// call; the step loop below only enqueues the kernels of the library (no host synchronisation, no host read-back), so a
// binder -- jax.ffi or the torch host -- sees the sweep as a single asynchronous operation on `stream`.
// Exchange proposals of two equal spins can never be accepted (metropolis.py:314-316) and are not evaluated when the
// tensor-core tower serves the model (moved chains compacted on the device, batch size read on the device).
#include "common.cuh"

namespace qtx {
struct SweepLayout {
  size_t new_spins, moved, rank, cspins, count, sig_new, ex_new, fwd, total;
  size_t fwd_bytes;
};
static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }
static SweepLayout sweep_layout(int mdt, int64_t ns, int nblocks, int channels, int lx, int ly, int kh, int kw) {
  SweepLayout L;
  const size_t N = (size_t)lx * ly;
  size_t off = 0;
  L.new_spins = off; off += al256((size_t)ns * N);
  L.moved = off; off += al256((size_t)ns);
  L.rank = off; off += al256((size_t)ns * 4);
  L.cspins = off; off += al256((size_t)ns * N);
  L.count = off; off += 256;
  L.sig_new = off; off += al256((size_t)ns * 8);
  L.ex_new = off; off += al256((size_t)ns * 8);
  L.fwd_bytes = qtx_resconv_workspace_size(mdt, ns, nblocks, channels, lx, ly, kh, kw, 0);
  L.fwd = off; off += al256(L.fwd_bytes);
  L.total = off + 256;
  return L;
}
}  // namespace qtx

using namespace qtx;

extern "C" size_t qtx_resconv_sweep_workspace_size(int model_dtype, int64_t ns, int nblocks, int channels, int lx, int ly,
                                                   int kh, int kw) {
  if (ns <= 0 || lx <= 0 || ly <= 0) return 0;
  const SweepLayout L = sweep_layout(model_dtype, ns, nblocks, channels, lx, ly, kh, kw);
  return L.fwd_bytes == 0 ? 0 : L.total;
}

extern "C" int qtx_resconv_sweep(int model_dtype, const void* params, int nblocks, int channels, int lx, int ly, int kh,
                                 int kw, int final_act, int8_t* spins, int64_t ns, int nsweeps, int kind,
                                 const int32_t* nbr_table, int max_nb, int hop, double reweight, uint64_t seed,
                                 uint64_t step0, uint64_t chain0, double* significand, double* exponent,
                                 int32_t* naccept, void* workspace, size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(params && spins && significand && exponent && workspace && ns > 0 && nsweeps >= 0 && lx > 0 && ly > 0,
              QTX_ERR_INVALID, "qtx_resconv_sweep: bad argument");
  QTX_REQUIRE(kind == QTX_LOCAL_FLIP || (kind == QTX_SPIN_EXCHANGE && nbr_table && max_nb > 0), QTX_ERR_INVALID,
              "qtx_resconv_sweep: exchange proposals need the neighbour table");
  const SweepLayout L = sweep_layout(model_dtype, ns, nblocks, channels, lx, ly, kh, kw);
  QTX_REQUIRE(L.fwd_bytes > 0 && workspace_bytes >= L.total, QTX_ERR_INVALID, "qtx_resconv_sweep: workspace too small");
  const int N = lx * ly;
  char* base = (char*)al256((size_t)workspace);
  int8_t* new_spins = (int8_t*)(base + L.new_spins);
  uint8_t* moved = (uint8_t*)(base + L.moved);
  int32_t* rank = (int32_t*)(base + L.rank);
  int8_t* cspins = (int8_t*)(base + L.cspins);
  int64_t* count = (int64_t*)(base + L.count);
  double* sig_new = (double*)(base + L.sig_new);
  double* ex_new = (double*)(base + L.ex_new);
  void* fwd = base + L.fwd;
  cudaStream_t st = (cudaStream_t)stream;
  if (naccept) QTX_CUDA(cudaMemsetAsync(naccept, 0, (size_t)ns * sizeof(int32_t), st));
  // psi of the current chains (Samples.psi is carried through the sweep)
  int rc = qtx_resconv_forward(model_dtype, params, nblocks, channels, lx, ly, kh, kw, final_act, spins, ns, significand,
                               exponent, fwd, L.fwd_bytes, stream);
  if (rc) return rc;
  const bool compact = kind == QTX_SPIN_EXCHANGE &&
                       qtx_resconv_tc_available(model_dtype, channels, lx, ly, kh, kw) != 0;
  if (compact) QTX_CUDA(cudaMemsetAsync(cspins, 0, (size_t)ns * N, st));
  for (int t = 0; t < nsweeps; ++t) {
    rc = qtx_metropolis_propose(kind, spins, ns, N, nbr_table, max_nb, hop, nullptr, nullptr, seed, step0 + t, chain0,
                                new_spins, moved, stream);
    if (rc) return rc;
    if (compact) {
      rc = qtx_compact_moved(moved, new_spins, ns, N, rank, cspins, count, stream);
      if (rc) return rc;
      rc = qtx_resconv_forward_n(model_dtype, params, nblocks, channels, lx, ly, kh, kw, final_act, 0, cspins, ns, count,
                                 sig_new, ex_new, fwd, L.fwd_bytes, stream);
      if (rc) return rc;
      rc = qtx_metropolis_accept_compact(spins, new_spins, moved, rank, ns, N, significand, exponent, sig_new, ex_new, 0,
                                         reweight, nullptr, seed, step0 + t, chain0, naccept, nullptr, stream);
    } else {
      rc = qtx_resconv_forward(model_dtype, params, nblocks, channels, lx, ly, kh, kw, final_act, new_spins, ns, sig_new,
                               ex_new, fwd, L.fwd_bytes, stream);
      if (rc) return rc;
      rc = qtx_metropolis_accept(spins, new_spins, moved, ns, N, significand, exponent, sig_new, ex_new, reweight, nullptr,
                                 seed, step0 + t, chain0, naccept, nullptr, stream);
    }
    if (rc) return rc;
  }
  return QTX_OK;
}
