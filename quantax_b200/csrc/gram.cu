// T = A A^T entry point (quantax/optimizer/solver.py:139): dispatch between the tcgen05 int8-sliced
// tensor-core kernel (gram_tc.cu) and the FP64 FMA cross-check path (gram_fma.cu).
#include "common.cuh"

namespace qtx {
int gram_fma(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, double* Tout, int accum, cudaStream_t st);
size_t gram_tc_workspace(int dtype, int64_t ns, int64_t np, int nslices);
int gram_tc(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* Tout, int accum,
            void* ws, size_t ws_bytes, cudaStream_t st);
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
