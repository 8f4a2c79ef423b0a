// Connected-configuration enumeration for spin Hamiltonians (generic-model path of Operator.Oloc):
// bit-exact mirror of _apply_off_diag + _get_conn (quantax/operator/operator.py:96-165) without
// ever materialising the [ns, nconn_raw, N] candidate tensor.  Integer / byte work, HBM bound:
// spins are staged per CTA in shared memory, compaction is an in-CTA ballot scan, and the
// connected configurations are written as coalesced rows.
#include "common.cuh"

namespace qtx {

struct TermEval {
  double c;
  bool nonnan;  // all '+'/'-' applications valid
  int nfl;
};

// applies term t to the configuration in shared memory `sp` (read-only); returns coefficient/validity
__device__ __forceinline__ TermEval eval_term(const int8_t* sp, double c, uint2 st, uint32_t op4) {
  int site[4] = {(int)(st.x & 0xffff), (int)(st.x >> 16), (int)(st.y & 0xffff), (int)(st.y >> 16)};
  TermEval r;
  r.nonnan = true;
  r.nfl = 0;
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    int op = (op4 >> (8 * k)) & 0xff;
    if (op == QTX_OP_NONE || op == QTX_OP_I) continue;
    int sk = sp[site[k]];
    if (op == QTX_OP_Z) {
      c = c * sk / 2;
    } else {
      if (op == QTX_OP_X) c = c / 2;
      else if (op == QTX_OP_P) r.nonnan = r.nonnan && (sk < 0);
      else r.nonnan = r.nonnan && (sk > 0);
      ++r.nfl;
    }
  }
  r.c = c;
  return r;
}

// value of site j after applying term (op codes) to sp: '+' sets +1, '-' sets -1, 'x' flips
__device__ __forceinline__ int8_t applied_spin(const int8_t* sp, int j, uint2 st, uint32_t op4) {
  int site[4] = {(int)(st.x & 0xffff), (int)(st.x >> 16), (int)(st.y & 0xffff), (int)(st.y >> 16)};
  int8_t v = sp[j];
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    int op = (op4 >> (8 * k)) & 0xff;
    if (site[k] != j) continue;
    if (op == QTX_OP_X) v = -v;
    else if (op == QTX_OP_P) v = 1;
    else if (op == QTX_OP_M) v = -1;
  }
  return v;
}

__global__ void __launch_bounds__(128) conn_count_kernel(const int8_t* __restrict__ spins, int N,
                                                         const double* __restrict__ coef,
                                                         const uint2* __restrict__ sites,
                                                         const uint32_t* __restrict__ ops, int nterms, int nflips_sel,
                                                         int32_t* __restrict__ nonnan_out,
                                                         int32_t* __restrict__ valid_out) {
  extern __shared__ int8_t sp[];
  __shared__ int red[2][4];
  const int64_t s = blockIdx.x;
  for (int j = threadIdx.x; j < N; j += blockDim.x) sp[j] = spins[s * N + j];
  __syncthreads();
  int nn = 0, nv = 0;
  for (int t = threadIdx.x; t < nterms; t += blockDim.x) {
    TermEval e = eval_term(sp, coef[t], sites[t], ops[t]);
    if (e.nfl == 0 || (nflips_sel > 0 && e.nfl != nflips_sel)) continue;
    nn += e.nonnan;
    nv += e.nonnan && fabs(e.c) > 1e-8;
  }
  nn = warp_sum(nn);
  nv = warp_sum(nv);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = nn;
    red[1][threadIdx.x >> 5] = nv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (nonnan_out) nonnan_out[s] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    if (valid_out) valid_out[s] = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  }
}

__global__ void __launch_bounds__(128) conn_fill_kernel(const int8_t* __restrict__ spins, int N,
                                                        const double* __restrict__ coef,
                                                        const uint2* __restrict__ sites,
                                                        const uint32_t* __restrict__ ops, int nterms, int nflips_sel,
                                                        const int64_t* __restrict__ offsets, int64_t conn_size,
                                                        int32_t* __restrict__ segment_out,
                                                        int32_t* __restrict__ conn_idx_out, double* __restrict__ H_out,
                                                        int8_t* __restrict__ s_conn_out) {
  extern __shared__ int8_t sp[];
  __shared__ int warp_cnt[4];
  __shared__ int list[128];
  const int64_t s = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < N; j += blockDim.x) sp[j] = spins[s * N + j];
  __syncthreads();
  int64_t base = offsets[s];
  for (int t0 = 0; t0 < nterms; t0 += blockDim.x) {
    int t = t0 + threadIdx.x;
    bool valid = false;
    double c = 0.0;
    if (t < nterms) {
      TermEval e = eval_term(sp, coef[t], sites[t], ops[t]);
      valid = e.nfl > 0 && (nflips_sel <= 0 || e.nfl == nflips_sel) && e.nonnan && fabs(e.c) > 1e-8;
      c = e.c;
    }
    uint32_t bal = __ballot_sync(FULL, valid);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      if (w < warp) before += warp_cnt[w];
      total += warp_cnt[w];
    }
    int my = before + __popc(bal & ((1u << lane) - 1u));
    if (valid) {
      int64_t pos = base + my;
      if (pos < conn_size) {
        segment_out[pos] = (int32_t)s;
        conn_idx_out[pos] = t;
        H_out[pos] = c;
      }
      list[my] = t;
    }
    __syncthreads();
    if (s_conn_out) {
      for (int e = warp; e < total; e += 4) {
        int64_t pos = base + e;
        if (pos >= conn_size) break;
        int tt = list[e];
        uint2 st = sites[tt];
        uint32_t op4 = ops[tt];
        int8_t* row = s_conn_out + pos * N;
        for (int j = lane; j < N; j += 32) row[j] = applied_spin(sp, j, st, op4);
      }
    }
    base += total;
    __syncthreads();
  }
}

// padding entries [total, conn_size): segment -1, conn -1, H 0, s_conn = candidate (last sample, last term)
__global__ void conn_pad_kernel(const int8_t* __restrict__ spins, int64_t ns, int N, const uint2* __restrict__ sites,
                                const uint32_t* __restrict__ ops, int nterms, const int64_t* __restrict__ total,
                                int64_t conn_size, int32_t* __restrict__ segment_out,
                                int32_t* __restrict__ conn_idx_out, double* __restrict__ H_out,
                                int8_t* __restrict__ s_conn_out) {
  const int64_t tot = *total;
  const int8_t* sp = spins + (ns - 1) * N;
  uint2 st = sites[nterms - 1];
  uint32_t op4 = ops[nterms - 1];
  for (int64_t pos = tot + blockIdx.x; pos < conn_size; pos += gridDim.x) {
    if (threadIdx.x == 0) {
      segment_out[pos] = -1;
      conn_idx_out[pos] = -1;
      H_out[pos] = 0.0;
    }
    if (s_conn_out)
      for (int j = threadIdx.x; j < N; j += blockDim.x) s_conn_out[pos * N + j] = applied_spin(sp, j, st, op4);
  }
}

__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ in, int64_t n,
                                                              int64_t* __restrict__ out, int64_t* __restrict__ total) {
  __shared__ int64_t part[1024];
  const int tid = threadIdx.x;
  int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
  int64_t a = tid * chunk, b = a + chunk < n ? a + chunk : n;
  int64_t sum = 0;
  for (int64_t i = a; i < b; ++i) sum += in[i];
  part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < blockDim.x; o <<= 1) {
    int64_t v = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int64_t run = part[tid] - sum;
  for (int64_t i = a; i < b; ++i) {
    out[i] = run;
    run += in[i];
  }
  if (tid == blockDim.x - 1 && total) *total = part[tid];
}

__global__ void __launch_bounds__(128) apply_diag_kernel(const int8_t* __restrict__ spins, int N,
                                                         const double* __restrict__ coef,
                                                         const uint2* __restrict__ sites,
                                                         const uint32_t* __restrict__ ops, int nterms,
                                                         double* __restrict__ diag_out) {
  extern __shared__ int8_t sp[];
  __shared__ double red[4];
  const int64_t s = blockIdx.x;
  for (int j = threadIdx.x; j < N; j += blockDim.x) sp[j] = spins[s * N + j];
  __syncthreads();
  double acc = 0.0;
  for (int t = threadIdx.x; t < nterms; t += blockDim.x) {
    TermEval e = eval_term(sp, coef[t], sites[t], ops[t]);
    if (e.nfl == 0) acc += e.c;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) diag_out[s] = (red[0] + red[1]) + (red[2] + red[3]);
}

// Eloc[seg] += H * (m'/m[seg]) * exp(e' - e[seg]); entries of one sample are contiguous, so a warp
// first combines equal segments (match_any) and issues one atomic per distinct segment.
__global__ void oloc_reduce_kernel(const int32_t* __restrict__ segment, const double* __restrict__ H,
                                   const double* __restrict__ mult_conn, const double* __restrict__ expo_conn,
                                   int64_t nconn, const double* __restrict__ mult, const double* __restrict__ expo,
                                   int64_t ns, double* __restrict__ eloc) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int seg = -1;
  double v = 0.0;
  if (c < nconn) {
    seg = segment[c];
    if (seg >= 0 && seg < ns) v = H[c] * ((mult_conn[c] / mult[seg]) * exp(expo_conn[c] - expo[seg]));
    else seg = -1;
  }
  const int lane = threadIdx.x & 31;
  uint32_t peers = __match_any_sync(FULL, seg);
  int leader = __ffs(peers) - 1;
  double sum = 0.0;
  for (uint32_t m = peers; m; m &= m - 1) {
    int src = __ffs(m) - 1;
    sum += __shfl_sync(peers, v, src);
  }
  if (lane == leader && seg >= 0) atomicAdd(eloc + seg, sum);
}

}  // namespace qtx

using namespace qtx;

extern "C" int qtx_conn_count(const int8_t* spins, int64_t ns, int N, const double* term_coef,
                              const uint16_t* term_sites, const uint8_t* term_ops, int nterms, int nflips_sel,
                              int32_t* nonnan_out, int32_t* valid_out, qtx_stream_t stream) {
  QTX_REQUIRE(spins && N > 0 && ns >= 0 && nterms >= 0, QTX_ERR_INVALID, "qtx_conn_count: bad argument");
  QTX_REQUIRE(nterms == 0 || (term_coef && term_sites && term_ops), QTX_ERR_INVALID, "qtx_conn_count: null terms");
  if (ns == 0) return QTX_OK;
  conn_count_kernel<<<(unsigned)ns, 128, (N + 15) / 16 * 16, (cudaStream_t)stream>>>(
      spins, N, term_coef, (const uint2*)term_sites, (const uint32_t*)term_ops, nterms, nflips_sel, nonnan_out,
      valid_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, int64_t* total_out,
                                      qtx_stream_t stream) {
  QTX_REQUIRE(in && out && n >= 0, QTX_ERR_INVALID, "qtx_exclusive_scan_i32: bad argument");
  exclusive_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(in, n, out, total_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_conn_fill(const int8_t* spins, int64_t ns, int N, const double* term_coef,
                             const uint16_t* term_sites, const uint8_t* term_ops, int nterms, int nflips_sel,
                             const int64_t* offsets, const int64_t* total, int64_t conn_size, int32_t* segment_out,
                             int32_t* conn_idx_out, double* H_out, int8_t* s_conn_out, qtx_stream_t stream) {
  QTX_REQUIRE(spins && offsets && total && segment_out && conn_idx_out && H_out && N > 0 && ns > 0 && nterms > 0,
              QTX_ERR_INVALID, "qtx_conn_fill: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  conn_fill_kernel<<<(unsigned)ns, 128, (N + 15) / 16 * 16, st>>>(
      spins, N, term_coef, (const uint2*)term_sites, (const uint32_t*)term_ops, nterms, nflips_sel, offsets,
      conn_size, segment_out, conn_idx_out, H_out, s_conn_out);
  QTX_LAUNCH_CHECK();
  conn_pad_kernel<<<4 * num_sms(), 128, 0, st>>>(spins, ns, N, (const uint2*)term_sites, (const uint32_t*)term_ops,
                                                 nterms, total, conn_size, segment_out, conn_idx_out, H_out,
                                                 s_conn_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_apply_diag(const int8_t* spins, int64_t ns, int N, const double* term_coef,
                              const uint16_t* term_sites, const uint8_t* term_ops, int nterms, double* diag_out,
                              qtx_stream_t stream) {
  QTX_REQUIRE(spins && diag_out && N > 0 && ns >= 0, QTX_ERR_INVALID, "qtx_apply_diag: bad argument");
  if (ns == 0) return QTX_OK;
  apply_diag_kernel<<<(unsigned)ns, 128, (N + 15) / 16 * 16, (cudaStream_t)stream>>>(
      spins, N, term_coef, (const uint2*)term_sites, (const uint32_t*)term_ops, nterms, diag_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_oloc_reduce(const int32_t* segment, const double* H, const double* mult_conn,
                               const double* expo_conn, int64_t nconn, const double* mult, const double* expo,
                               int64_t ns, double* eloc_inout, qtx_stream_t stream) {
  QTX_REQUIRE(segment && H && mult_conn && expo_conn && mult && expo && eloc_inout, QTX_ERR_INVALID,
              "qtx_oloc_reduce: bad argument");
  if (nconn == 0) return QTX_OK;
  oloc_reduce_kernel<<<(unsigned)((nconn + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      segment, H, mult_conn, expo_conn, nconn, mult, expo, ns, eloc_inout);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
