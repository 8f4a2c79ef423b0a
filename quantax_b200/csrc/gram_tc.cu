// T = A A^T on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM,
// operands staged by TMA) with float64-grade accuracy.
//
// tcgen05 has no FP64 MMA kind, and the reference computes the MinSR Gram matrix in float64
// (quantax/optimizer/solver.py:139 on the float64 Jacobian of variational.py:491).  We therefore
// use an error-free transformation (Ozaki scheme):
//   1. every row r of A is scaled by a power of two 2^-e_r so that |v| < 1, and cut into `s`
//      signed digits  v = q_1 2^-6 + q_2 2^-13 + ... + q_s 2^-(7s-1) + O(2^-7s), |q_a| <= 64 (int8);
//   2. the int8 digit matrices are multiplied pairwise on the tensor cores with EXACT int32
//      accumulation:  S_ab[i,j] = sum_k q_a[i,k] q_b[j,k].  Pairs with the same level d = a + b
//      share the weight 2^(-7d+2) and therefore one TMEM accumulator; pairs with d > s + 1 are
//      below the truncation error and skipped (s (s+1) / 2 products instead of s^2);
//   3. the epilogue recombines the levels in float64:
//      T[i,j] = 2^(e_i+e_j) sum_d 2^(-7d+2) S_d[i,j]  -- every term is exact, the sum is rounded.
// Only tiles on or below the diagonal are computed; the epilogue mirrors them.
//
// Kernel structure (one CTA per SM, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA
// issuer (single thread) + TMEM allocator, warps 2-5 = epilogue (TMEM -> registers -> float64 ->
// global).  A tile is 128 x 128 outputs; TMEM holds up to four level accumulators of 128 columns
// (all 512 columns), so `s` digits need ceil(s / 4) passes over K.  One pipeline stage holds one
// 32-element K block of all needed digit slices of both row panels (2 * s * 4 KB).
#include <stdlib.h>

#include <vector>

#include "gram_tc_common.cuh"

namespace qtx {

// ---------------------------------------------------------------------------------------------
// slicing kernel: one CTA per (padded) row
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gram_split_kernel(const T* __restrict__ A, int64_t ns, int64_t np, int64_t ld,
                                                         int64_t k0, int64_t kc, int64_t kc_pad, int64_t ns_pad,
                                                         int nslices, int8_t* __restrict__ Q,
                                                         double* __restrict__ rowscale, int first_chunk, int have_scale) {
  __shared__ double red[8];
  __shared__ double s_inv;
  const int64_t r = blockIdx.x;
  const int tid = threadIdx.x;
  double inv = 0.0;
  if (have_scale) {
    // several K chunks: the scales were computed once by this kernel's scale-only launch (kc_pad == 0); re-reading the
    // full row for every chunk cost 10 x 17 GB at config E
    const double sc = r < ns ? rowscale[r] : 0.0;
    inv = (sc > 0.0 && isfinite(sc)) ? 1.0 / sc : 0.0;  // power of two: exact
  } else if (r < ns) {
    // row scale from the FULL row (all K chunks share it)
    double mx = 0.0;
    const T* row = A + r * ld;
    for (int64_t k = tid; k < np; k += blockDim.x) mx = fmax(mx, fabs((double)row[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
      double sc = 1.0;
      if (!isfinite(mx)) {
        sc = nan("");
        s_inv = 0.0;
      } else if (mx > 0.0) {
        int e;
        frexp(mx, &e);  // mx = m 2^e, m in [0.5, 1)  ->  |x| 2^-e < 1
        sc = scalbn(1.0, e);
        s_inv = scalbn(1.0, -e);
      } else {
        s_inv = 0.0;
      }
      if (first_chunk) rowscale[r] = sc;
    }
    __syncthreads();
    inv = s_inv;
  } else if (tid == 0 && first_chunk) {
    rowscale[r] = 0.0;
  }
  // 4 consecutive k per thread -> one coalesced 4-byte store per slice (a warp writes 128 B)
  const T* row = A + (r < ns ? r : 0) * ld;
  for (int64_t kb = (int64_t)tid * 4; kb < kc_pad; kb += (int64_t)blockDim.x * 4) {
    uint32_t w[kMaxSlices];
#pragma unroll
    for (int a = 0; a < kMaxSlices; ++a) w[a] = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int64_t k = k0 + kb + e;
      double v = (r < ns && kb + e < kc && k < np) ? (double)row[k] * inv : 0.0;
      double res = v * 64.0;
#pragma unroll
      for (int a = 0; a < kMaxSlices; ++a) {
        if (a < nslices) {
          double qa = rint(res);
          w[a] |= ((uint32_t)(int)qa & 0xffu) << (8 * e);
          res = (res - qa) * 128.0;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < kMaxSlices; ++a)
      if (a < nslices) {
        // K-blocked layout [slice][kblock][row][32 B]: one TMA box (128 rows x 32 B) is a contiguous 4 KB run
        const int64_t nkb = kc_pad / kBK;
        *reinterpret_cast<uint32_t*>(Q + (((int64_t)a * nkb + kb / kBK) * ns_pad + r) * kBK + (kb % kBK)) = w[a];
      }
  }
}

// upper triangle <- lower triangle (the CTA-pair kernel writes j <= i only): 32 x 32 blocks through shared memory,
// both sides coalesced
__global__ void __launch_bounds__(256) gram_mirror_kernel(double* __restrict__ T, int64_t ns) {
  __shared__ double tile[32][33];
  const int nbk = (int)((ns + 31) / 32);
  const int64_t nblocks = (int64_t)nbk * (nbk + 1) / 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    int bi = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((int64_t)(bi + 1) * (bi + 2) / 2 <= b) ++bi;
    while ((int64_t)bi * (bi + 1) / 2 > b) --bi;
    const int bj = (int)(b - (int64_t)bi * (bi + 1) / 2);
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int64_t i = (int64_t)bi * 32 + r, j = (int64_t)bj * 32 + tx;
      tile[r][tx] = (i < ns && j < ns) ? T[i * ns + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int64_t j = (int64_t)bj * 32 + r, i = (int64_t)bi * 32 + tx;  // T[j][i] = T[i][j] for j < i
      if (i < ns && j < ns && j < i) T[j * ns + i] = tile[tx][r];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ void tile_from_index(int t, int& I, int& J) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((i + 1) * (i + 2) / 2 <= t) ++i;
  while (i * (i + 1) / 2 > t) --i;
  I = i;
  J = t - i * (i + 1) / 2;
}

// One K block of one pass: all digit pairs (a, b) with level d = a + b in the pass, fully unrolled at
// compile time so that the single issuing thread spends ~3 instructions per MMA (the descriptors of a
// stage differ from the stage-0 descriptors only by a constant added to the low word).
template <int S, int PASS>
__device__ __forceinline__ void issue_kblock(uint32_t tmem_base, uint32_t desc_lo, uint32_t desc_hi, uint32_t acc_first) {
  constexpr int d_lo = 2 + PASS * kLevelsPerPass;
  constexpr int d_hi = (d_lo + kLevelsPerPass - 1 < S + 1) ? d_lo + kLevelsPerPass - 1 : S + 1;
#pragma unroll
  for (int d = d_lo; d <= d_hi; ++d) {
    const int a_lo = (d - S > 1) ? d - S : 1, a_hi = (d - 1 < S) ? d - 1 : S;
#pragma unroll
    for (int a = a_lo; a <= a_hi; ++a) {
      const int b = d - a;
      const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo + (uint32_t)(a - 1) * (kSliceBytes >> 4));
      const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo + (uint32_t)(S + b - 1) * (kSliceBytes >> 4));
      umma_i8(tmem_base + (uint32_t)(d - d_lo) * kTile, da, db, kIdescI8, a == a_lo ? acc_first : 1u);
    }
  }
}

template <int S>
__global__ void __launch_bounds__(kThreads, 1) gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, GramTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t stage_bytes = 2u * S * kSliceBytes;
  constexpr int npasses = (S + kLevelsPerPass - 1) / kLevelsPerPass;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);  // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int ntiles = p.nb * (p.nb + 1) / 2;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int I, J;
        tile_from_index(t, I, J);
#pragma unroll
        for (int pass = 0; pass < npasses; ++pass) {
          const int d_hi = min(2 + (pass + 1) * kLevelsPerPass - 1, S + 1);
          const int nsl = min(S, d_hi - 1);
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            unsigned char* base = smem + (size_t)stage * stage_bytes;
            mbar_expect_tx(full_bar + stage, 2u * nsl * kSliceBytes);
#pragma unroll
            for (int a = 0; a < S; ++a) {
              if (a < nsl) {
                tma_load_3d(base + (size_t)a * kSliceBytes, &tmap, full_bar + stage, 0, I * kTile, a * p.nkb + kb);
                tma_load_3d(base + (size_t)(S + a) * kSliceBytes, &tmap, full_bar + stage, 0, J * kTile, a * p.nkb + kb);
              }
            }
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, tphase = 0;
      const uint64_t desc0 = make_desc_sw32(smem_u32(smem));
      const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
#pragma unroll
        for (int pass = 0; pass < npasses; ++pass) {
          mbar_wait(tmem_empty, tphase ^ 1);  // epilogue has drained the accumulators
          tc_fence_after();
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            const uint32_t desc_lo = desc_lo0 + stage * (stage_bytes >> 4);
            const uint32_t acc_first = kb > 0 ? 1u : 0u;
            if (pass == 0) issue_kblock<S, 0>(tmem_base, desc_lo, desc_hi, acc_first);
            else issue_kblock<S, (npasses > 1 ? 1 : 0)>(tmem_base, desc_lo, desc_hi, acc_first);
            umma_commit(empty_bar + stage);  // frees the smem stage once these MMAs have read it
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
          umma_commit(tmem_full);  // accumulators complete
          tphase ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lanes 32*(warp%4) .. +31 =====
    const int lg = warp & 3;
    uint32_t tphase = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int I, J;
      tile_from_index(t, I, J);
      const int64_t row = (int64_t)I * kTile + lg * 32 + lane;
      const double rs_i = (row < p.ns) ? p.rowscale[row] : 0.0;
      for (int pass = 0; pass < npasses; ++pass) {
        const int d_lo = 2 + pass * kLevelsPerPass;
        const int d_hi = min(d_lo + kLevelsPerPass - 1, S + 1);
        mbar_wait(tmem_full, tphase);
        tc_fence_after();
        const bool add = (pass > 0) || (p.accum != 0);
        for (int c0 = 0; c0 < kTile; c0 += 16) {
          double sum[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) sum[e] = 0.0;
          for (int d = d_hi; d >= d_lo; --d) {  // smallest weights first
            int32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(d - d_lo) * kTile + c0, v);
            tmem_ld_wait();
            const double w = scalbn(1.0, -7 * d + 2);
#pragma unroll
            for (int e = 0; e < 16; ++e) sum[e] += w * (double)v[e];
          }
          if (row < p.ns) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int64_t col = (int64_t)J * kTile + c0 + e;
              if (col < p.ns && (I != J || col <= row)) {
                const double val = sum[e] * (rs_i * p.rowscale[col]);
                double* p1 = p.T + row * p.ns + col;
                *p1 = add ? *p1 + val : val;
                if (col != row) {
                  double* p2 = p.T + col * p.ns + row;
                  *p2 = add ? *p2 + val : val;
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty);
        tphase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int S>
static int launch_gram_tc(const CUtensorMap& tmap, const GramTcParams& p, int grid, size_t smem, cudaStream_t st) {
  QTX_CUDA(cudaFuncSetAttribute(gram_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_tc_kernel<S><<<grid, kThreads, smem, st>>>(tmap, p);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

int gram_tc2_launch(int s, const CUtensorMap& tmA, const CUtensorMap& tmB, const GramTcParams& p, int grid, size_t smem,
                    cudaStream_t st);
int gram_tc2_launch_push(int s, const CUtensorMap& tmA, const CUtensorMap& tmB, const GramTcParams& p, int grid,
                         size_t smem, cudaStream_t st, int nranks, int rank, double* const* slots);

static bool use_2cta() {
  const char* e = getenv("QTX_GRAM_2CTA");
  return e ? atoi(e) != 0 : true;
}

static int default_slices(int dtype) { return dtype == QTX_F64 ? 7 : 4; }  // s = 7: 1.7e-15 of |a_i||a_j| (cuBLAS DGEMM: 7.9e-15)

static void gram_tc_sizes(int dtype, int64_t ns, int64_t np, int nslices, int& s, int64_t& ns_pad, int64_t& kc,
                          int64_t& kc_pad) {
  s = nslices > 0 ? nslices : default_slices(dtype);
  ns_pad = (ns + 2 * kTile - 1) / (2 * kTile) * (2 * kTile);  // multiple of 256 (CTA-pair tiles)
  // exact int32 accumulation: up to s pairs per level, |q| <= 64  ->  K * s * 4096 < 2^31
  int64_t kmax = ((int64_t)1 << 31) / (4096 * (int64_t)s) - 1;
  kmax = kmax / 64 * 64;
  // equal K chunks: the kernels always run over kc_pad columns (zero padded), so a short last chunk cost as much as a
  // full one (config E shard on 8 GPUs: 130 944 columns = 104 832 + 26 112, both 265 ms)
  const int64_t nchunks = (np + kmax - 1) / kmax;
  kc = ((np + nchunks - 1) / nchunks + 63) / 64 * 64;
  if (kc > kmax) kc = kmax;
  kc_pad = (kc + 63) / 64 * 64;
}

size_t gram_tc_workspace(int dtype, int64_t ns, int64_t np, int nslices) {
  int s;
  int64_t ns_pad, kc, kc_pad;
  gram_tc_sizes(dtype, ns, np, nslices, s, ns_pad, kc, kc_pad);
  const size_t nb2 = (size_t)(ns_pad / (2 * kTile));
  return (size_t)s * ns_pad * kc_pad + (size_t)ns_pad * sizeof(double) + nb2 * (nb2 + 1) * sizeof(int) + 256 + 1024;
}

// push_slots != nullptr: fused Gram + exchange (qtx_gram_push) -- the last K chunk runs the PUSH kernel, which also
// stores every finished tile into push_slots[q] (q != push_rank), the staging slots of the peers.
static int gram_tc_impl(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* Tout,
                        int accum, void* ws, size_t ws_bytes, cudaStream_t st, int push_nranks, int push_rank,
                        double* const* push_slots) {
  int s;
  int64_t ns_pad, kc, kc_pad;
  gram_tc_sizes(dtype, ns, np, nslices, s, ns_pad, kc, kc_pad);
  QTX_REQUIRE(s >= 1 && s <= kMaxSlices, QTX_ERR_INVALID, "qtx_gram: nslices must be in 1..%d", kMaxSlices);
  QTX_REQUIRE(ws && ws_bytes >= gram_tc_workspace(dtype, ns, np, nslices), QTX_ERR_INVALID,
              "qtx_gram: workspace too small");
  QTX_REQUIRE(ns_pad / kTile < 30000, QTX_ERR_UNSUPPORTED, "qtx_gram: ns too large");
  EncodeTiledFn encode = get_encode_fn();
  QTX_REQUIRE(encode != nullptr, QTX_ERR_CUDA, "qtx_gram: cuTensorMapEncodeTiled is unavailable");
  double* rowscale = reinterpret_cast<double*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  const int nb2_map = (int)(ns_pad / (2 * kTile));
  int* tile_map = reinterpret_cast<int*>(((uintptr_t)(rowscale + ns_pad) + 255) & ~(uintptr_t)255);
  int8_t* Q = reinterpret_cast<int8_t*>(((uintptr_t)(tile_map + (size_t)nb2_map * (nb2_map + 1)) + 255) & ~(uintptr_t)255);

  CUtensorMap tmap;
  // dims: (byte in K block, row, slice * nkb + kblock)
  cuuint64_t gdim[3] = {(cuuint64_t)kBK, (cuuint64_t)ns_pad, (cuuint64_t)s * (cuuint64_t)(kc_pad / kBK)};
  cuuint64_t gstride[2] = {(cuuint64_t)kBK, (cuuint64_t)kBK * (cuuint64_t)ns_pad};
  cuuint32_t box[3] = {kBK, kTile, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, Q, gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QTX_REQUIRE(cr == CUDA_SUCCESS, QTX_ERR_CUDA, "qtx_gram: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const bool pair = use_2cta();
  QTX_REQUIRE(pair || !push_slots, QTX_ERR_UNSUPPORTED, "qtx_gram_push needs the CTA-pair kernel (QTX_GRAM_2CTA=1)");
  CUtensorMap tmapB;  // 64-row boxes: each CTA of a pair stages half of the J panel
  if (pair) {
    cuuint32_t boxB[3] = {kBK, kTile / 2, 1};
    cr = encode(&tmapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, Q, gdim, gstride, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    QTX_REQUIRE(cr == CUDA_SUCCESS, QTX_ERR_CUDA, "qtx_gram: cuTensorMapEncodeTiled (B) failed (%d)", (int)cr);
  }
  const uint32_t stage_bytes = pair ? (uint32_t)s * (kSliceBytes + kSliceBytes / 2) : 2u * s * kSliceBytes;
  const size_t epi_bytes = pair ? (size_t)4 * 32 * 17 * sizeof(double) + 64 : 0;  // transpose blocks of the pair kernel's epilogue
  int stages = (int)((220 * 1024 - epi_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (const char* e = getenv("QTX_GRAM_STAGES")) {
    int v = atoi(e);
    if (v >= 2 && v <= stages) stages = v;
  }
  QTX_REQUIRE(stages >= 2, QTX_ERR_UNSUPPORTED, "qtx_gram: pipeline does not fit in shared memory");
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align*/ + (2 * stages + 2) * 8 + 16 + epi_bytes;

  GramTcParams p;
  p.tile_map = nullptr;
  if (pair) {
    // supertile order of the lower-triangular (256-row block, 128-column block) tiles: bands of 6 row blocks, walked in
    // chunks of 12 column blocks (see tile2_from_index)
    static thread_local std::vector<int> order;
    order.clear();
    const int GI = 6, GJ = 12;
    for (int b = 0; b < nb2_map; b += GI) {
      const int i1 = b + GI < nb2_map ? b + GI : nb2_map;
      const int jmax = 2 * (i1 - 1) + 1;
      for (int jc = 0; jc <= jmax; jc += GJ)
        for (int i = b; i < i1; ++i)
          for (int j = jc; j < jc + GJ && j <= 2 * i + 1; ++j) order.push_back((i << 16) | j);
    }
    QTX_REQUIRE((int)order.size() == nb2_map * (nb2_map + 1), QTX_ERR_INVALID, "qtx_gram: tile order");
    // measured (tools/gram_order_probe.py): no gain at config B (7.6 vs 7.6 ms) or the config E slice (47.1 vs 47.7 ms) and a
    // loss on the 16384-row shard of config E on 8 GPUs (362 vs 333 ms): with 5 or 7 digits the kernel is bound by the
    // L2 -> shared-memory stream of its own CTA pair, which the order does not change.  Opt-in: QTX_GRAM_SUPERTILE=1
    const char* e = getenv("QTX_GRAM_SUPERTILE");
    if (e && e[0] == '1') {
      QTX_CUDA(cudaMemcpyAsync(tile_map, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      p.tile_map = tile_map;
    }
  }
  p.nb = (int)(ns_pad / kTile);
  p.nslices = s;
  p.stages = stages;
  p.ns = ns;
  p.rowscale = rowscale;
  p.T = Tout;
  int grid;
  if (pair) {
    const int nb2 = p.nb / 2, ntiles2 = nb2 * (nb2 + 1), maxcl = num_sms() / 2;
    grid = 2 * (ntiles2 < maxcl ? ntiles2 : maxcl);
  } else {
    const int ntiles = p.nb * (p.nb + 1) / 2;
    grid = ntiles < num_sms() ? ntiles : num_sms();
  }
  int chunk = 0;
  const int multi = np > kc ? 1 : 0;
  for (int64_t k0 = 0; k0 < np; k0 += kc, ++chunk) {
    const int64_t kcur = (np - k0) < kc ? (np - k0) : kc;
    const int64_t kcur_pad = (kcur + 63) / 64 * 64;
    QTX_REQUIRE(kcur_pad == kc_pad || chunk > 0, QTX_ERR_INVALID, "qtx_gram: internal chunk error");
    if (chunk == 0 && multi) {  // scale-only launch: no columns to cut (kc_pad = 0), writes rowscale
      if (dtype == QTX_F64)
        gram_split_kernel<double><<<(unsigned)ns_pad, 256, 0, st>>>((const double*)A, ns, np, ld, 0, 0, 0, ns_pad, s, Q,
                                                                    rowscale, 1, 0);
      else
        gram_split_kernel<float><<<(unsigned)ns_pad, 256, 0, st>>>((const float*)A, ns, np, ld, 0, 0, 0, ns_pad, s, Q,
                                                                   rowscale, 1, 0);
      QTX_LAUNCH_CHECK();
    }
    if (dtype == QTX_F64)
      gram_split_kernel<double><<<(unsigned)ns_pad, 256, 0, st>>>((const double*)A, ns, np, ld, k0, kcur, kc_pad,
                                                                  ns_pad, s, Q, rowscale, chunk == 0 && !multi, multi);
    else
      gram_split_kernel<float><<<(unsigned)ns_pad, 256, 0, st>>>((const float*)A, ns, np, ld, k0, kcur, kc_pad, ns_pad,
                                                                 s, Q, rowscale, chunk == 0 && !multi, multi);
    QTX_LAUNCH_CHECK();
    p.nkb = (int)(kc_pad / kBK);
    p.accum = (accum != 0 || chunk > 0) ? 1 : 0;
    int rc = QTX_OK;
    if (pair) {
      if (push_slots && k0 + kc >= np)
        rc = gram_tc2_launch_push(s, tmap, tmapB, p, grid, smem, st, push_nranks, push_rank, push_slots);
      else
        rc = gram_tc2_launch(s, tmap, tmapB, p, grid, smem, st);
      if (rc) return rc;
      continue;
    }
    switch (s) {
      case 1: rc = launch_gram_tc<1>(tmap, p, grid, smem, st); break;
      case 2: rc = launch_gram_tc<2>(tmap, p, grid, smem, st); break;
      case 3: rc = launch_gram_tc<3>(tmap, p, grid, smem, st); break;
      case 4: rc = launch_gram_tc<4>(tmap, p, grid, smem, st); break;
      case 5: rc = launch_gram_tc<5>(tmap, p, grid, smem, st); break;
      case 6: rc = launch_gram_tc<6>(tmap, p, grid, smem, st); break;
      case 7: rc = launch_gram_tc<7>(tmap, p, grid, smem, st); break;
      default: rc = launch_gram_tc<8>(tmap, p, grid, smem, st); break;
    }
    if (rc) return rc;
  }
  if (pair && !push_slots) {  // the fused path's reduce kernel writes both triangles itself
    gram_mirror_kernel<<<8 * num_sms(), 256, 0, st>>>(Tout, ns);
    QTX_LAUNCH_CHECK();
  }
  return QTX_OK;
}

int gram_tc(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* Tout, int accum,
            void* ws, size_t ws_bytes, cudaStream_t st) {
  return gram_tc_impl(dtype, A, ns, np, ld, nslices, Tout, accum, ws, ws_bytes, st, 0, 0, nullptr);
}

int gram_tc_push(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices, double* Tout, void* ws,
                 size_t ws_bytes, cudaStream_t st, int nranks, int rank, double* const* slots) {
  return gram_tc_impl(dtype, A, ns, np, ld, nslices, Tout, 0, ws, ws_bytes, st, nranks, rank, slots);
}

}  // namespace qtx
