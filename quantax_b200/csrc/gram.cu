// T = A A^T entry point (quantax/optimizer/solver.py:139): dispatch between the tcgen05 int8-sliced
// tensor-core kernel (gram_tc.cu) and the FP64 FMA cross-check path (gram_fma.cu).
#include "common.cuh"

namespace qtx {
int gram_fma(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, double* Tout, int accum, cudaStream_t st);
size_t gram_tc_workspace(int dtype, int64_t ns, int64_t np, int nslices);
int gram_tc(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* Tout, int accum,
            void* ws, size_t ws_bytes, cudaStream_t st);
int gram_tc_push(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* Tout, void* ws,
                 size_t ws_bytes, cudaStream_t st, int nranks, int rank, double* const* slots);
}  // namespace qtx

using namespace qtx;

extern "C" size_t qtx_gram_workspace_size(int dtype, int64_t ns, int64_t np, int nslices) {
  if (nslices < 0) return 256;
  return gram_tc_workspace(dtype, ns, np, nslices);
}

extern "C" int qtx_gram(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* T_out,
                        int T_accum, void* workspace, size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(A && T_out && ns > 0 && np > 0 && ld >= np, QTX_ERR_INVALID, "qtx_gram: bad argument");
  QTX_REQUIRE(dtype == QTX_F32 || dtype == QTX_F64, QTX_ERR_INVALID, "qtx_gram: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (nslices < 0) return gram_fma(dtype, A, ns, np, ld, T_out, T_accum, st);
  return gram_tc(dtype, A, ns, np, ld, nslices, T_out, T_accum, workspace, workspace_bytes, st);
}

// Fused Gram + exchange: the partial T of this rank is computed as by qtx_gram and every finished tile is also
// stored (lower triangle) into peer_slots[q], q != rank -- device pointers into the peers' staging areas, mapped with
// qtx_peer_open.  Follow with qtx_peer_signal and qtx_gram_reduce (peer.cu).
extern "C" int qtx_gram_push(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* T_out,
                             int nranks, int rank, void* const* peer_slots, void* workspace, size_t workspace_bytes,
                             qtx_stream_t stream) {
  QTX_REQUIRE(A && T_out && ns > 0 && np > 0 && ld >= np, QTX_ERR_INVALID, "qtx_gram_push: bad argument");
  QTX_REQUIRE(dtype == QTX_F32 || dtype == QTX_F64, QTX_ERR_INVALID, "qtx_gram_push: bad dtype %d", dtype);
  QTX_REQUIRE(nslices >= 0, QTX_ERR_UNSUPPORTED, "qtx_gram_push: the FMA cross-check kernel has no push variant");
  QTX_REQUIRE(peer_slots && nranks >= 1 && nranks <= QTX_MAX_PEERS && rank >= 0 && rank < nranks, QTX_ERR_INVALID,
              "qtx_gram_push: bad rank layout");
  double* slots[QTX_MAX_PEERS] = {};
  for (int q = 0; q < nranks; ++q) {
    QTX_REQUIRE(q == rank || peer_slots[q], QTX_ERR_INVALID, "qtx_gram_push: null slot for rank %d", q);
    slots[q] = (double*)peer_slots[q];
  }
  return gram_tc_push(dtype, A, ns, np, ld, nslices, T_out, workspace, workspace_bytes, (cudaStream_t)stream, nranks,
                      rank, slots);
}
