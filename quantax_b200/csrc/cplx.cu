// Complex-amplitude variants of the small reductions of the VMC step, used by states with real
// parameters and complex output (VS_TYPE.real_to_complex, quantax/state/variational.py:256-257):
//   Oloc reduction with complex psi ratios          operator.py:168-184
//   state-level symmetry projection of complex psi  symmetry.py:386-392, variational.py:438-491
//   Ebar of complex local energies, stacked [Re; Im] for the real solver  sr.py:99-104,180-195
// psi travels as ScaleArray(significand complex128 [n] interleaved (re, im), exponent float64 [n]).
// All kernels are bandwidth bound.
#include "common.cuh"

namespace qtx {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// Eloc[seg] += H * (m'/m[seg]) * exp(e' - e[seg]); one atomic pair per distinct segment of a warp
__global__ void oloc_reduce_cplx_kernel(const int32_t* __restrict__ segment, const double* __restrict__ H,
                                        const double2* __restrict__ mult_conn, const double* __restrict__ expo_conn,
                                        int64_t nconn, const double2* __restrict__ mult, const double* __restrict__ expo,
                                        int64_t ns, double2* __restrict__ eloc) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int seg = -1;
  double2 v = make_double2(0.0, 0.0);
  if (c < nconn) {
    seg = segment[c];
    if (seg >= 0 && seg < ns) {
      const double2 r = cdiv(mult_conn[c], mult[seg]);
      const double f = H[c] * exp(expo_conn[c] - expo[seg]);
      v = make_double2(r.x * f, r.y * f);
    } else {
      seg = -1;
    }
  }
  const int lane = threadIdx.x & 31;
  uint32_t peers = __match_any_sync(FULL, seg);
  int leader = __ffs(peers) - 1;
  double sr = 0.0, si = 0.0;
  for (uint32_t m = peers; m; m &= m - 1) {
    int src = __ffs(m) - 1;
    sr += __shfl_sync(peers, v.x, src);
    si += __shfl_sync(peers, v.y, src);
  }
  if (lane == leader && seg >= 0) {
    atomicAdd(&eloc[seg].x, sr);
    atomicAdd(&eloc[seg].y, si);
  }
}

// psi = sum_g w_g psi_g for ScaleArray images with complex significands (real weights); coef_g = w_g psi_g / psi
__global__ void symm_combine_cplx_kernel(const double2* __restrict__ mult, const double* __restrict__ expo, int64_t ns,
                                         int nsymm, const double* __restrict__ w, double2* __restrict__ mult_out,
                                         double* __restrict__ expo_out, double2* __restrict__ coef_out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= ns) return;
  const double2* m = mult + s * nsymm;
  const double* e = expo + s * nsymm;
  double emax = -INFINITY;
  for (int g = lane; g < nsymm; g += 32) emax = fmax(emax, e[g]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(FULL, emax, o));
  double br = 0.0, bi = 0.0;
  for (int g = lane; g < nsymm; g += 32) {
    const double r = ((e[g] != emax) ? exp(e[g] - emax) : 1.0) * w[g];
    br += m[g].x * r;
    bi += m[g].y * r;
  }
  br = warp_sum(br);
  bi = warp_sum(bi);
  const double2 b = make_double2(br, bi);
  if (coef_out)
    for (int g = lane; g < nsymm; g += 32) {
      const double r = ((e[g] != emax) ? exp(e[g] - emax) : 1.0) * w[g];
      coef_out[s * nsymm + g] = cdiv(make_double2(m[g].x * r, m[g].y * r), b);
    }
  if (lane == 0) {
    mult_out[s] = b;
    expo_out[s] = emax;
  }
}

// Stacked real Jacobians: J holds Re O(T_g s) in rows [0, R) and Im O(T_g s) in rows [j_im_off, j_im_off + R),
// R = ns * nsymm; out receives Re O(s) in rows [0, ns) and Im O(s) in rows [o_im_off, o_im_off + ns):
//   O(s) = sum_g coef[s, g] * O(T_g s)   (complex coef)
template <typename T>
__global__ void __launch_bounds__(256) weighted_rowsum_cplx_kernel(const T* __restrict__ J, int64_t ldj, int64_t j_im_off,
                                                                   const double2* __restrict__ coef, int nsymm, int64_t np,
                                                                   T* __restrict__ out, int64_t ldo, int64_t o_im_off) {
  const int64_t s = blockIdx.y;
  const double2* c = coef + s * nsymm;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    double ar = 0.0, ai = 0.0;
    for (int g = 0; g < nsymm; ++g) {
      const double jr = (double)J[(s * nsymm + g) * ldj + k], ji = (double)J[(j_im_off + s * nsymm + g) * ldj + k];
      ar += c[g].x * jr - c[g].y * ji;
      ai += c[g].x * ji + c[g].y * jr;
    }
    out[s * ldo + k] = (T)ar;
    out[(o_im_off + s) * ldo + k] = (T)ai;
  }
}

// single-CTA statistics of complex local energies (sr.py:180-195): stats = (Re <E rw>, <|E - <E rw>|^2 rw>),
// ebar_stacked[s] = Re, ebar_stacked[ns + s] = Im of (E - <E>) sqrt(rw / ns)
__global__ void __launch_bounds__(1024) ebar_cplx_kernel(const double2* __restrict__ eloc, const double* __restrict__ rw,
                                                         int64_t ns, double* __restrict__ ebar, int64_t im_off,
                                                         double* __restrict__ stats) {
  __shared__ double red[4][32];
  __shared__ double bc[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  for (int64_t s = tid; s < ns; s += blockDim.x) {
    const double2 e = eloc[s];
    const double r = rw ? rw[s] : 1.0;
    a0 += e.x * r; a1 += e.y * r; b0 += e.x; b1 += e.y;
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); b0 = warp_sum(b0); b1 = warp_sum(b1);
  if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = b0; red[3][warp] = b1; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    a0 = lane < nw ? red[0][lane] : 0.0; a1 = lane < nw ? red[1][lane] : 0.0;
    b0 = lane < nw ? red[2][lane] : 0.0; b1 = lane < nw ? red[3][lane] : 0.0;
    a0 = warp_sum(a0); a1 = warp_sum(a1); b0 = warp_sum(b0); b1 = warp_sum(b1);
    if (lane == 0) { bc[0] = a0 / (double)ns; bc[1] = a1 / (double)ns; bc[2] = b0 / (double)ns; bc[3] = b1 / (double)ns; }
  }
  __syncthreads();
  double v = 0.0;
  for (int64_t s = tid; s < ns; s += blockDim.x) {
    const double2 e = eloc[s];
    const double r = rw ? rw[s] : 1.0;
    const double dr = e.x - bc[0], di = e.y - bc[1];
    v += (dr * dr + di * di) * r;
    if (ebar) {
      const double f = sqrt(r / (double)ns);
      ebar[s] = (e.x - bc[2]) * f;
      ebar[im_off + s] = (e.y - bc[3]) * f;
    }
  }
  __syncthreads();
  v = warp_sum(v);
  if (lane == 0) red[0][warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? red[0][lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0 && stats) { stats[0] = bc[0]; stats[1] = v / (double)ns; }
  }
}

__global__ void real_to_cplx_kernel(const double* __restrict__ x, int64_t n, double2* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = make_double2(x[i], 0.0);
}

}  // namespace qtx

using namespace qtx;

extern "C" int qtx_oloc_reduce_cplx(const int32_t* segment, const double* H, const double* mult_conn_c128,
                                    const double* expo_conn, int64_t nconn, const double* mult_c128, const double* expo,
                                    int64_t ns, double* eloc_c128_inout, qtx_stream_t stream) {
  QTX_REQUIRE(segment && H && mult_conn_c128 && expo_conn && mult_c128 && expo && eloc_c128_inout, QTX_ERR_INVALID,
              "qtx_oloc_reduce_cplx: bad argument");
  if (nconn == 0) return QTX_OK;
  oloc_reduce_cplx_kernel<<<(unsigned)((nconn + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      segment, H, (const double2*)mult_conn_c128, expo_conn, nconn, (const double2*)mult_c128, expo, ns,
      (double2*)eloc_c128_inout);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_symm_combine_cplx(const double* mult_c128, const double* expo, int64_t ns, int nsymm,
                                     const double* weights, double* mult_c128_out, double* expo_out, double* coef_c128_out,
                                     qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(mult_c128 && expo && weights && mult_c128_out && expo_out && nsymm > 0, QTX_ERR_INVALID,
              "qtx_symm_combine_cplx: bad argument");
  symm_combine_cplx_kernel<<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const double2*)mult_c128, expo, ns, nsymm, weights, (double2*)mult_c128_out, expo_out, (double2*)coef_c128_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_weighted_rowsum_cplx(int dtype, const void* J, int64_t ldj, int64_t j_im_row_offset,
                                        const double* coef_c128, int64_t ns, int nsymm, int64_t np, void* out, int64_t ldo,
                                        int64_t out_im_row_offset, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(J && coef_c128 && out && nsymm > 0 && np > 0 && ldj >= np && ldo >= np && ns <= 65535, QTX_ERR_INVALID,
              "qtx_weighted_rowsum_cplx: bad argument");
  QTX_REQUIRE(j_im_row_offset >= ns * nsymm && out_im_row_offset >= ns, QTX_ERR_INVALID,
              "qtx_weighted_rowsum_cplx: imaginary blocks overlap the real ones");
  unsigned gx = (unsigned)((np + 255) / 256);
  if (gx > 256) gx = 256;
  dim3 grid(gx, (unsigned)ns);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTX_F64)
    weighted_rowsum_cplx_kernel<double><<<grid, 256, 0, st>>>((const double*)J, ldj, j_im_row_offset,
                                                             (const double2*)coef_c128, nsymm, np, (double*)out, ldo,
                                                             out_im_row_offset);
  else if (dtype == QTX_F32)
    weighted_rowsum_cplx_kernel<float><<<grid, 256, 0, st>>>((const float*)J, ldj, j_im_row_offset,
                                                            (const double2*)coef_c128, nsymm, np, (float*)out, ldo,
                                                            out_im_row_offset);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_weighted_rowsum_cplx: bad dtype %d", dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_ebar_cplx(const double* eloc_c128, const double* rw, int64_t ns, double* ebar_stacked_out,
                             int64_t im_offset, double* stats_out, qtx_stream_t stream) {
  QTX_REQUIRE(eloc_c128 && ns > 0 && (ebar_stacked_out == nullptr || im_offset >= ns), QTX_ERR_INVALID,
              "qtx_ebar_cplx: bad argument");
  ebar_cplx_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const double2*)eloc_c128, rw, ns, ebar_stacked_out, im_offset,
                                                         stats_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_real_to_cplx(const double* x, int64_t n, double* out_c128, qtx_stream_t stream) {
  if (n == 0) return QTX_OK;
  QTX_REQUIRE(x && out_c128, QTX_ERR_INVALID, "qtx_real_to_cplx: bad argument");
  unsigned g = (unsigned)((n + 255) / 256);
  if (g > 1024) g = 1024;
  real_to_cplx_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(x, n, (double2*)out_c128);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
