// Complex-symmetric LDL^T factorisation and triangular solves for the shifted systems (T - z I) x = b of the soft
// pseudo-inverse (pinv_rational.cu; quantax/optimizer/solver.py:94-111,142-146 compute the same y = f(T) b from eigh).
// Own kernels, FP64 FMA pipe, no library call.
//
// M = T - z I is complex SYMMETRIC (not Hermitian) and its field of values is the segment [-z, lambda_max - z], which
// stays at distance Im z > 0 from the origin: e^{i phi} M has a positive definite Hermitian part for a suitable phi, so
// the factorisation M = L D L^T (L unit lower triangular, D diagonal, both complex) needs NO pivoting.  Only the lower
// triangle of M is read and written: L below the diagonal, D on it.
//
// Right-looking, block size NB = 64, three launches per block column:
//   zldlt_diag_kernel   : one CTA factorises the NB x NB diagonal block in shared memory;
//   zldlt_panel_kernel  : CTA per 64-row tile below it: W = A21 L11^-T (column sweep in shared memory), L21 = W D^-1;
//                         L21 goes back into M, W and L21 also into k-major scratch panels WT / LT [NB][n];
//   zldlt_update_kernel : CTA per 64 x 64 tile of the lower triangle of the trailing matrix: C -= W L21^T, operands
//                         staged k-major in shared memory (conflict-free 16-byte loads), 4 x 4 complex register tile
//                         per thread (64 DFMA per k step and thread).
// Solves L D L^T x = r run as two persistent wavefront kernels (ztrsv_kernel<false/true>): CTA per 64-row block,
// block rows are handed out in dependency order through an atomic ticket, a CTA consumes the solution blocks it
// depends on as soon as their flag is published, solves its diagonal block and publishes its own.
#ifdef QTX_HOST_EMULATION
#include "cuda_emu.h"
#include "cuda_emu_host.h"
#else
#include <cuComplex.h>

#include "common.cuh"
#define QTX_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define QTX_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_raw[]; T* name = reinterpret_cast<T*>(name##_raw)
#endif
#include "zldlt.cuh"

namespace qtx {

#ifdef QTX_HOST_EMULATION
constexpr int kNB = 8, kR = 2;  // small blocks: a few std::threads per CTA, several block columns at n ~ 40
#else
constexpr int kNB = 64, kR = 4;
#endif
constexpr int kTB = kNB / kR;          // threads per tile edge
constexpr int kThreads = kTB * kTB;    // 256
constexpr int kPad = kNB + 1;          // row pitch (complex numbers) of the row-major shared-memory tiles
constexpr int kKH = kNB / 2;           // k extent staged per pass of the update kernel
constexpr int kTPR = kThreads / kNB;   // threads cooperating on one row / column of a 64 x 64 block product

typedef cuDoubleComplex cplx;

__device__ __forceinline__ cplx cmake(double re, double im) { return make_cuDoubleComplex(re, im); }
__device__ __forceinline__ cplx cmul_(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void cfms(cplx& acc, cplx a, cplx b) {  // acc -= a * b
  acc.x -= a.x * b.x;
  acc.x += a.y * b.y;
  acc.y -= a.x * b.y;
  acc.y -= a.y * b.x;
}
__device__ __forceinline__ void cfma_(cplx& acc, cplx a, cplx b) {  // acc += a * b
  acc.x += a.x * b.x;
  acc.x -= a.y * b.y;
  acc.y += a.x * b.y;
  acc.y += a.y * b.x;
}
__device__ __forceinline__ cplx crecip(cplx d) {  // 1 / d, scaled against overflow of |d|^2
  const double s = fmax(fabs(d.x), fabs(d.y));
  if (s == 0.0) return cmake(0.0, 0.0);
  const double a = d.x / s, b = d.y / s;
  const double q = 1.0 / ((a * a + b * b) * s);
  return cmake(a * q, -b * q);
}

// ---- factorisation ----------------------------------------------------------------------------------------------
// diagonal block [k0, k0 + nb): in place L11 (strictly lower) and D1 (diagonal).  A zero pivot sets info = its 1-based
// index (first one wins) and is replaced by 1 so that the factorisation continues with finite numbers.
__global__ void __launch_bounds__(kThreads) zldlt_diag_kernel(cplx* __restrict__ M, int64_t n, int64_t k0, int nb,
                                                             int32_t* __restrict__ info) {
  QTX_DYN_SMEM(cplx, sm);
  cplx* A = sm;                  // [kNB][kPad]
  cplx* wv = sm + kNB * kPad;    // [kNB]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < nb * nb; idx += kThreads) {
    const int i = idx / nb, j = idx % nb;
    if (j <= i) A[i * kPad + j] = M[(k0 + i) * n + k0 + j];
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    cplx d = A[k * kPad + k];
    if (d.x == 0.0 && d.y == 0.0) {
      if (tid == 0 && info[0] == 0) info[0] = (int32_t)(k0 + k + 1);
      d = cmake(1.0, 0.0);
    }
    const cplx inv = crecip(d);
    const int m = nb - k - 1;
    if (tid < m) {
      const int i = k + 1 + tid;
      const cplx w = A[i * kPad + k];
      wv[i] = w;
      A[i * kPad + k] = cmul_(w, inv);
    }
    __syncthreads();  // also orders the read of A[k][k] above against the write-back of the patched pivot below
    if (tid == 0) A[k * kPad + k] = d;
    for (int idx = tid; idx < m * m; idx += kThreads) {
      const int i = k + 1 + idx / m, j = k + 1 + idx % m;
      if (j <= i) cfms(A[i * kPad + j], A[i * kPad + k], wv[j]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nb * nb; idx += kThreads) {
    const int i = idx / nb, j = idx % nb;
    if (j <= i) M[(k0 + i) * n + k0 + j] = A[i * kPad + j];
  }
}

// rows [k0 + nb, n) of block column [k0, k0 + nb): W = A21 L11^-T, L21 = W D1^-1
__global__ void __launch_bounds__(kThreads) zldlt_panel_kernel(cplx* __restrict__ M, int64_t n, int64_t k0, int nb,
                                                              cplx* __restrict__ WT, cplx* __restrict__ LT) {
  QTX_DYN_SMEM(cplx, sm);
  cplx* L11 = sm;               // [kNB][kPad]: strictly lower = L, diagonal = D
  cplx* X = sm + kNB * kPad;    // [kNB][kPad]: row tile of A21 -> W
  const int tid = threadIdx.x;
  const int64_t i0 = k0 + nb + (int64_t)blockIdx.x * kNB;
  const int nr = (int)((n - i0) < kNB ? (n - i0) : kNB);
  for (int idx = tid; idx < nb * nb; idx += kThreads) {
    const int i = idx / nb, j = idx % nb;
    if (j <= i) L11[i * kPad + j] = M[(k0 + i) * n + k0 + j];
  }
  for (int idx = tid; idx < nr * nb; idx += kThreads) {
    const int r = idx / nb, c = idx % nb;
    X[r * kPad + c] = M[(i0 + r) * n + k0 + c];
  }
  __syncthreads();
  for (int c = 0; c + 1 < nb; ++c) {  // column c of X is final (= W[:, c]); eliminate it from the columns behind it
    const int rest = nb - c - 1;
    for (int idx = tid; idx < nr * rest; idx += kThreads) {
      const int r = idx % nr, c2 = c + 1 + idx / nr;
      cfms(X[r * kPad + c2], X[r * kPad + c], L11[c2 * kPad + c]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nr * nb; idx += kThreads) {  // L21 into M: c fastest (row-major M)
    const int r = idx / nb, c = idx % nb;
    M[(i0 + r) * n + k0 + c] = cmul_(X[r * kPad + c], crecip(L11[c * kPad + c]));
  }
  for (int idx = tid; idx < nr * nb; idx += kThreads) {  // k-major panels: r fastest
    const int r = idx % nr, c = idx / nr;
    const cplx w = X[r * kPad + c];
    WT[(int64_t)c * n + i0 + r] = w;
    LT[(int64_t)c * n + i0 + r] = cmul_(w, crecip(L11[c * kPad + c]));
  }
}

// trailing update: tile (I, J), J <= I, of the lower triangle behind block column [k0, k0 + nb):  C -= W L21^T
__global__ void __launch_bounds__(kThreads, 2) zldlt_update_kernel(cplx* __restrict__ M, int64_t n, int64_t k0, int nb,
                                                                  const cplx* __restrict__ WT,
                                                                  const cplx* __restrict__ LT) {
  QTX_DYN_SMEM(cplx, sm);
  cplx* Ws = sm;                 // [kKH][kNB]  (k-major: one k = 64 consecutive rows)
  cplx* Ls = sm + kKH * kNB;     // [kKH][kNB]
  const int tid = threadIdx.x, tx = tid % kTB, ty = tid / kTB;
  // linear block index -> (I, J) with J <= I
  const unsigned b = blockIdx.x;
  int I = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((unsigned)I * (unsigned)(I + 1) / 2u > b) --I;
  while ((unsigned)(I + 1) * (unsigned)(I + 2) / 2u <= b) ++I;
  const int J = (int)(b - (unsigned)I * (unsigned)(I + 1) / 2u);
  const int64_t t0 = k0 + nb;
  const int64_t i0 = t0 + (int64_t)I * kNB, j0 = t0 + (int64_t)J * kNB;
  const int nr = (int)((n - i0) < kNB ? (n - i0) : kNB), nc = (int)((n - j0) < kNB ? (n - j0) : kNB);
  cplx acc[kR][kR];
#pragma unroll
  for (int a = 0; a < kR; ++a)
#pragma unroll
    for (int c = 0; c < kR; ++c) acc[a][c] = cmake(0.0, 0.0);
  for (int kh = 0; kh < nb; kh += kKH) {
    const int kn = (nb - kh) < kKH ? (nb - kh) : kKH;
    for (int idx = tid; idx < kKH * kNB; idx += kThreads) {
      const int k = idx / kNB, r = idx % kNB;
      const bool kin = k < kn;
      Ws[idx] = (kin && r < nr) ? WT[(int64_t)(kh + k) * n + i0 + r] : cmake(0.0, 0.0);
      Ls[idx] = (kin && r < nc) ? LT[(int64_t)(kh + k) * n + j0 + r] : cmake(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kKH; ++k) {
      cplx av[kR], bv[kR];
#pragma unroll
      for (int a = 0; a < kR; ++a) av[a] = Ws[k * kNB + ty + kTB * a];
#pragma unroll
      for (int c = 0; c < kR; ++c) bv[c] = Ls[k * kNB + tx + kTB * c];
#pragma unroll
      for (int a = 0; a < kR; ++a)
#pragma unroll
        for (int c = 0; c < kR; ++c) cfma_(acc[a][c], av[a], bv[c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < kR; ++a) {
    const int r = ty + kTB * a;
    if (r >= nr) continue;
    const int64_t i = i0 + r;
#pragma unroll
    for (int c = 0; c < kR; ++c) {
      const int cc = tx + kTB * c;
      const int64_t j = j0 + cc;
      if (cc < nc && j <= i) {
        cplx v = M[i * n + j];
        v.x -= acc[a][c].x;
        v.y -= acc[a][c].y;
        M[i * n + j] = v;
      }
    }
  }
}

// inverse of the unit lower triangular diagonal blocks, all blocks in parallel (CTA per block): Gauss-Jordan on
// [L | I] in shared memory; row k of the inverse is final when step k starts.  invL [nblk][kNB][kNB], row-major,
// zero above the diagonal and outside a ragged last block.
__global__ void __launch_bounds__(kThreads) zldlt_diaginv_kernel(const cplx* __restrict__ M, int64_t n,
                                                                cplx* __restrict__ invL) {
  QTX_DYN_SMEM(cplx, sm);
  cplx* L = sm;                 // [kNB][kPad]
  cplx* X = sm + kNB * kPad;    // [kNB][kPad]
  const int tid = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.x * kNB;
  const int nr = (int)((n - i0) < kNB ? (n - i0) : kNB);
  for (int idx = tid; idx < kNB * kNB; idx += kThreads) {
    const int i = idx / kNB, j = idx % kNB;
    L[i * kPad + j] = (i < nr && j < i) ? M[(i0 + i) * n + i0 + j] : cmake(0.0, 0.0);
    X[i * kPad + j] = cmake(i == j ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  for (int k = 0; k + 1 < nr; ++k) {  // rows r > k: X[r][0..k] -= L[r][k] X[k][0..k]
    const int rows = nr - k - 1, cols = k + 1;
    for (int idx = tid; idx < rows * cols; idx += kThreads) {
      const int r = k + 1 + idx / cols, c = idx % cols;
      cfms(X[r * kPad + c], L[r * kPad + k], X[k * kPad + c]);
    }
    __syncthreads();
  }
  cplx* out = invL + (int64_t)blockIdx.x * kNB * kNB;
  for (int idx = tid; idx < kNB * kNB; idx += kThreads) out[idx] = X[(idx / kNB) * kPad + idx % kNB];
}

// ---- triangular solves ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_flag(const unsigned* p) { return *(const volatile unsigned*)p; }

// kBackward = false:  x <- L^-1 x          (block rows in increasing order)
// kBackward = true :  x <- L^-T D^-1 x     (block rows in decreasing order)
// sync[0] = ticket counter, sync[1 + b] = flag of block row b; both zeroed before the launch.
template <bool kBackward>
__global__ void __launch_bounds__(kThreads) ztrsv_kernel(const cplx* __restrict__ M, const cplx* __restrict__ invL,
                                                        int64_t n, cplx* x, unsigned* sync) {
  QTX_DYN_SMEM(cplx, sm);
  cplx* Li = sm;                        // [kNB][kPad] inverse of the diagonal block of L
  cplx* acc = sm + kNB * kPad;          // [kNB]
  cplx* xj = acc + kNB;                 // [kNB]
  cplx* part = xj + kNB;                // [kNB][kTPR]
  __shared__ unsigned ticket_s;
  const int tid = threadIdx.x;
  const int nblk = (int)((n + kNB - 1) / kNB);
  if (tid == 0) ticket_s = atomicAdd(sync, 1u);
  __syncthreads();
  const int b = kBackward ? nblk - 1 - (int)ticket_s : (int)ticket_s;
  const int64_t i0 = (int64_t)b * kNB;
  const int nr = (int)((n - i0) < kNB ? (n - i0) : kNB);
  for (int idx = tid; idx < kNB * kNB; idx += kThreads)
    Li[(idx / kNB) * kPad + idx % kNB] = invL[(int64_t)b * kNB * kNB + idx];
  if (tid < kNB) {
    cplx v = cmake(0.0, 0.0);
    if (tid < nr) {
      v = x[i0 + tid];
      if (kBackward) v = cmul_(v, crecip(M[(i0 + tid) * n + i0 + tid]));  // D^-1
    }
    acc[tid] = v;
  }
  const int line = tid / kTPR, p = tid % kTPR;     // forward: (row, part) -- kTPR threads share a row of the block
  const int col = tid % kNB, q = tid / kNB;        // backward: (column, part) -- coalesced along the rows of M
  cplx mine = cmake(0.0, 0.0);
  const int jbeg = kBackward ? nblk - 1 : 0, jend = b, jstep = kBackward ? -1 : 1;
  for (int j = jbeg; j != jend; j += jstep) {
    if (tid == 0) {
      unsigned long long spins = 0;
      while (ld_flag(sync + 1 + j) == 0u) {
        if (++spins > (1ull << 34)) {
#ifndef QTX_HOST_EMULATION
          __trap();  // a dependency that never arrives is a bug: fail instead of hanging the GPU
#endif
        }
      }
    }
    __syncthreads();
#ifndef QTX_HOST_EMULATION
    __threadfence();
#endif
    const int64_t j0 = (int64_t)j * kNB;
    const int nj = (int)((n - j0) < kNB ? (n - j0) : kNB);
    if (tid < kNB) {
#ifdef QTX_HOST_EMULATION
      xj[tid] = tid < nj ? x[j0 + tid] : cmake(0.0, 0.0);
#else
      cplx v = cmake(0.0, 0.0);
      if (tid < nj) {
        const double2 t = __ldcg(reinterpret_cast<const double2*>(x + j0 + tid));
        v = cmake(t.x, t.y);
      }
      xj[tid] = v;
#endif
    }
    __syncthreads();
    if (!kBackward) {
      if (line < nr) {
        const cplx* row = M + (i0 + line) * n + j0;
        for (int c = p; c < nj; c += kTPR) cfms(mine, row[c], xj[c]);
      }
    } else {
      if (col < nr) {
        for (int r = q; r < nj; r += kTPR) cfms(mine, M[(j0 + r) * n + i0 + col], xj[r]);
      }
    }
    __syncthreads();  // xj is rewritten by the next block
  }
  if (!kBackward) part[line * kTPR + p] = mine;
  else part[col * kTPR + q] = mine;
  __syncthreads();
  if (tid < kNB) {
    cplx v = acc[tid];
#pragma unroll
    for (int t = 0; t < kTPR; ++t) {
      v.x += part[tid * kTPR + t].x;
      v.y += part[tid * kTPR + t].y;
    }
    acc[tid] = v;
  }
  __syncthreads();
  // diagonal block: forward y = Linv acc, backward x = Linv^T acc (rows / columns beyond nr are zero in Linv)
  mine = cmake(0.0, 0.0);
  if (!kBackward) {
    for (int c = p; c < kNB; c += kTPR) cfma_(mine, Li[line * kPad + c], acc[c]);
    part[line * kTPR + p] = mine;
  } else {
    for (int r = q; r < kNB; r += kTPR) cfma_(mine, Li[r * kPad + col], acc[r]);
    part[col * kTPR + q] = mine;
  }
  __syncthreads();
  if (tid < nr) {
    cplx v = cmake(0.0, 0.0);
#pragma unroll
    for (int t = 0; t < kTPR; ++t) {
      v.x += part[tid * kTPR + t].x;
      v.y += part[tid * kTPR + t].y;
    }
    x[i0 + tid] = v;
  }
#ifndef QTX_HOST_EMULATION
  __threadfence();
#endif
  __syncthreads();
  if (tid == 0) {
#ifdef QTX_HOST_EMULATION
    sync[1 + b] = 1u;
#else
    atomicExch(sync + 1 + b, 1u);
#endif
  }
}

__global__ void zero_sync_kernel(unsigned* sync, int count) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) sync[i] = 0u;
}

// ---- host side --------------------------------------------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

constexpr size_t kSmemDiag = (size_t)(kNB * kPad + kNB) * sizeof(cplx);
constexpr size_t kSmemPanel = (size_t)(2 * kNB * kPad) * sizeof(cplx);
constexpr size_t kSmemUpdate = (size_t)(2 * kKH * kNB) * sizeof(cplx);
constexpr size_t kSmemTrsv = (size_t)(kNB * kPad + 2 * kNB + kNB * kTPR) * sizeof(cplx);

struct ZldltScratch {
  cplx *WT, *LT, *invL;
  unsigned* sync;  // [2][nblk + 1]: forward and backward sweep
};

static size_t panel_bytes(int64_t n) { return align256((size_t)kNB * (size_t)n * sizeof(cplx)); }
static size_t invl_bytes(int64_t n) { return align256((size_t)((n + kNB - 1) / kNB) * kNB * kNB * sizeof(cplx)); }

size_t zldlt_scratch_bytes(int64_t n) {
  const size_t nblk = (size_t)((n + kNB - 1) / kNB);
  return 2 * panel_bytes(n) + invl_bytes(n) + align256(2 * (nblk + 1) * sizeof(unsigned)) + 256;
}

static ZldltScratch carve(void* scratch, int64_t n) {
  char* base = (char*)align256((size_t)scratch);
  ZldltScratch s;
  s.WT = (cplx*)base;
  s.LT = (cplx*)(base + panel_bytes(n));
  s.invL = (cplx*)(base + 2 * panel_bytes(n));
  s.sync = (unsigned*)(base + 2 * panel_bytes(n) + invl_bytes(n));
  return s;
}

static int zldlt_prepare() {
#ifndef QTX_HOST_EMULATION
  static thread_local int prepared_device = -1;  // the attribute is per function and device
  int dev = 0;
  QTX_CUDA(cudaGetDevice(&dev));
  if (prepared_device != dev) {
    QTX_CUDA(cudaFuncSetAttribute(zldlt_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemDiag));
    QTX_CUDA(cudaFuncSetAttribute(zldlt_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPanel));
    QTX_CUDA(cudaFuncSetAttribute(zldlt_diaginv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPanel));
    QTX_CUDA(cudaFuncSetAttribute(zldlt_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemUpdate));
    QTX_CUDA(cudaFuncSetAttribute(ztrsv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTrsv));
    QTX_CUDA(cudaFuncSetAttribute(ztrsv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTrsv));
    prepared_device = dev;
  }
#endif
  return QTX_OK;
}

int zldlt_factor(cuDoubleComplex* M, int64_t n, void* scratch, int32_t* info, cudaStream_t st) {
  int rc = zldlt_prepare();
  if (rc) return rc;
  const ZldltScratch s = carve(scratch, n);
  for (int64_t k0 = 0; k0 < n; k0 += kNB) {
    const int nb = (int)((n - k0) < kNB ? (n - k0) : kNB);
    QTX_LAUNCH_SMEM(zldlt_diag_kernel, 1, kThreads, kSmemDiag, st, M, n, k0, nb, info);
    QTX_LAUNCH_CHECK();
    const int64_t rest = n - k0 - nb;
    if (rest <= 0) break;
    const unsigned tiles = (unsigned)((rest + kNB - 1) / kNB);
    QTX_LAUNCH_SMEM(zldlt_panel_kernel, tiles, kThreads, kSmemPanel, st, M, n, k0, nb, s.WT, s.LT);
    QTX_LAUNCH_CHECK();
    QTX_LAUNCH_SMEM(zldlt_update_kernel, tiles * (tiles + 1) / 2, kThreads, kSmemUpdate, st, M, n, k0, nb, s.WT, s.LT);
    QTX_LAUNCH_CHECK();
  }
  const unsigned nblk = (unsigned)((n + kNB - 1) / kNB);
  QTX_LAUNCH_SMEM(zldlt_diaginv_kernel, nblk, kThreads, kSmemPanel, st, M, n, s.invL);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

int zldlt_solve(const cuDoubleComplex* M, int64_t n, cuDoubleComplex* x, void* scratch, cudaStream_t st) {
  int rc = zldlt_prepare();
  if (rc) return rc;
  const ZldltScratch s = carve(scratch, n);
  const unsigned nblk = (unsigned)((n + kNB - 1) / kNB);
  QTX_LAUNCH_SMEM(zero_sync_kernel, 1, 256, 0, st, s.sync, (int)(2 * (nblk + 1)));
  QTX_LAUNCH_CHECK();
  QTX_LAUNCH_SMEM(ztrsv_kernel<false>, nblk, kThreads, kSmemTrsv, st, M, s.invL, n, x, s.sync);
  QTX_LAUNCH_CHECK();
  QTX_LAUNCH_SMEM(ztrsv_kernel<true>, nblk, kThreads, kSmemTrsv, st, M, s.invL, n, x, s.sync + nblk + 1);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

int zldlt_block_size() { return kNB; }

}  // namespace qtx
