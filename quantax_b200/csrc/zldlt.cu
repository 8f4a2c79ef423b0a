// Complex-symmetric LDL^T factorisation and triangular solves for the shifted systems (T - z I) x = b of the soft
// pseudo-inverse (pinv_rational.cu; quantax/optimizer/solver.py:94-111,142-146 compute the same y = f(T) b from eigh).
// Own kernels on the FP64 pipes (DMMA m8n8k4 for the trailing updates), no library call.
//
// M = T - z I is complex SYMMETRIC (not Hermitian) and its field of values is the segment [-z, lambda_max - z], which
// stays at distance Im z > 0 from the origin: e^{i phi} M has a positive definite Hermitian part for a suitable phi, so
// the factorisation M = L D L^T (L unit lower triangular, D diagonal, both complex) needs NO pivoting.  Only the lower
// triangle of M is read and written: L below the diagonal, D on it.
//
// Right-looking with look-ahead, block size NB = 64, ONE launch per block column (zldlt_step_kernel):
//   * every CTA owns a 64 x 64 tile of the lower triangle of the trailing matrix and applies C -= W L21^T of the
//     previous block column: operands staged as real / imaginary planes in shared memory, FP64 tensor-core MMAs
//     (mma.sync m8n8k4, four per complex 8 x 8 x 4 product), accumulators in registers;
//   * the CTA of tile (0, 0) then factorises its tile -- the NEXT diagonal block -- in registers (one barrier per
//     column) and publishes L11, D and 1 / D;
//   * the CTAs of the tiles (I, 0) below it wait for that flag and turn their updated tiles into the next panel:
//     W = A21 L11^-T by a column sweep in registers, L21 = W D^-1; L21 goes into M, W into a panel buffer.
//   Tiles are handed out through an atomic ticket, column tiles first, so the diagonal CTA is always running when
//   somebody waits for it; the rest of the trailing update hides the panel's latency.
// Solves L D L^T x = r run as two persistent wavefront kernels (ztrsv_kernel<false/true>): CTA per 64-row block,
// block rows handed out in dependency order through an atomic ticket, a CTA consumes the solution blocks it depends
// on as soon as their flag is published, multiplies by the inverse of its diagonal block (zldlt_diaginv_kernel,
// all blocks in parallel after the factorisation) and publishes its own.
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
constexpr int kNB = 16, kWR = 1, kWC = 2;  // small blocks, two emulated warps: several block columns at n ~ 40
#else
constexpr int kNB = 64, kWR = 2, kWC = 4;  // 8 warps, warp tile 32 x 16
#endif
constexpr int kThreads = 32 * kWR * kWC;
constexpr int kWTR = kNB / kWR, kWTC = kNB / kWC;  // warp tile
constexpr int kFI = kWTR / 8, kFJ = kWTC / 8;      // 8 x 8 accumulator fragments per warp tile
constexpr int kPad = kNB + 1;                      // row pitch (complex numbers) of the row-major shared-memory tiles
constexpr int kKC = kNB / 4;                       // k extent of one pipeline stage of the trailing update
constexpr int kKP = kKC + 4;                       // row pitch (complex numbers) of a stage: (16 kKP) mod 128 = 64 makes
                                                   // the 16-byte fragment loads of a quarter warp conflict-free
constexpr int kTPR = kThreads / kNB;               // threads cooperating on one row / column of a block product
static_assert(kKC % 4 == 0 && kWTR % 8 == 0 && kWTC % 8 == 0, "tile shapes");

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
// 1 / s for a positive finite double without the IEEE division sequence (it sits on the pivot-to-pivot critical path
// of the factorisation): exponent removed by bit manipulation, float reciprocal of the mantissa, three Newton steps
// (23 -> 46 -> 92 bits, the third one absorbs the rounding of the first two).
__device__ __forceinline__ double fast_recip_pos(double s) {
#ifdef QTX_HOST_EMULATION
  return 1.0 / s;
#else
  const int hi = __double2hiint(s);
  const int ex = ((hi >> 20) & 0x7ff) - 1023;               // unbiased exponent
  if (ex < -1000 || ex > 1000) return 1.0 / s;               // subnormal / huge: the slow exact path
  const double m = __hiloint2double(hi - (ex << 20), __double2loint(s));  // s 2^-ex in [1, 2)
  double r = (double)__frcp_rn((float)m);
  r = r * (2.0 - m * r);
  r = r * (2.0 - m * r);
  r = r * (2.0 - m * r);
  return __hiloint2double(__double2hiint(r) - (ex << 20), __double2loint(r));  // r 2^-ex
#endif
}
__device__ __forceinline__ cplx crecip(cplx d) {  // 1 / d = conj(d) / |d|^2
  const double s = d.x * d.x + d.y * d.y;
  if (!(s > 0.0)) return cmake(0.0, 0.0);
  if (s > 1e300 || s < 1e-300) {  // |d|^2 out of range: scale first
    const double m = fmax(fabs(d.x), fabs(d.y));
    const double a = d.x / m, b = d.y / m;
    const double q = 1.0 / ((a * a + b * b) * m);
    return cmake(a * q, -b * q);
  }
  const double r = fast_recip_pos(s);
  return cmake(d.x * r, -d.y * r);
}
// Hand-over between CTAs of one launch (CUTLASS semaphore pattern): the producer's threads write, __syncthreads(), then
// ONE thread stores the flag with release semantics at gpu scope (cumulative over the barrier); the consumer's thread
// 0 polls with acquire loads, __syncthreads(), and everybody reads the published data through L2 (ld.global.cg).
__device__ __forceinline__ unsigned ld_flag(const unsigned* p) {
#ifdef QTX_HOST_EMULATION
  return *(const volatile unsigned*)p;
#else
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ cplx ld_cg(const cplx* p) {  // a value another CTA of this launch has just published
#ifdef QTX_HOST_EMULATION
  return *p;
#else
  const double2 t = __ldcg(reinterpret_cast<const double2*>(p));
  return cmake(t.x, t.y);
#endif
}
__device__ __forceinline__ void spin_until_set(const unsigned* flag) {  // one thread; bounded: trap, never hang
  unsigned long long spins = 0;
  while (ld_flag(flag) == 0u) {
    if (++spins > (1ull << 34)) {
#ifndef QTX_HOST_EMULATION
      __trap();
#endif
    }
  }
}
__device__ __forceinline__ void publish(unsigned* flag) {  // by one thread, after the __syncthreads() behind the writes
#ifdef QTX_HOST_EMULATION
  *flag = 1u;
#else
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
#endif
}

// 16-byte asynchronous global -> shared copy (zero fill when !valid); groups are committed / waited per pipeline stage
__device__ __forceinline__ void cp_async16(cplx* smem_dst, const cplx* gmem_src, bool valid) {
#ifdef QTX_HOST_EMULATION
  *smem_dst = valid ? *gmem_src : cmake(0.0, 0.0);
#else
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef QTX_HOST_EMULATION
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
#ifndef QTX_HOST_EMULATION
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
#endif
}

// D (8 x 8) += A (8 x 4, row) B (4 x 8, col) in float64: lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and
// C[l / 4][2 (l % 4) + {0, 1}]  (PTX mma.sync.m8n8k4.f64)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
#ifdef QTX_HOST_EMULATION
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < 4; ++k) {
    const double ak = __shfl_sync(FULL, a, g * 4 + k);
    const double b0 = __shfl_sync(FULL, b, (2 * q) * 4 + k);
    const double b1 = __shfl_sync(FULL, b, (2 * q + 1) * 4 + k);
    s0 += ak * b0;
    s1 += ak * b1;
  }
  c0 += s0;
  c1 += s1;
#else
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
#endif
}

// ---- factorisation ----------------------------------------------------------------------------------------------
// A 64 x 64 complex tile lives in the registers of the CTA in the accumulator layout of the MMAs: thread (warp w,
// lane l) holds, for fragment (fi, fj) and e in {0, 1}, the element
//   row = (w / kWC) kWTR + 8 fi + l / 4,   col = (w % kWC) kWTC + 8 fj + 2 (l % 4) + e.
struct TilePos {
  int row0, col0;  // row of fragment row 0, column of fragment column 0, e = 0
  __device__ __forceinline__ int row(int fi) const { return row0 + 8 * fi; }
  __device__ __forceinline__ int col(int fj, int e) const { return col0 + 8 * fj + e; }
};
__device__ __forceinline__ TilePos tile_pos() {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  return {(w / kWC) * kWTR + (l >> 2), (w % kWC) * kWTC + 2 * (l & 3)};
}

// The pivot sweeps below run on a tile in SHARED memory with a compact loop body: one barrier-to-barrier step of the
// factorisation is a dependent chain of a few hundred cycles (measured: barrier 30, STS-BAR-LDS 90, reciprocal 90,
// DFMA 9), and an unrolled in-register version -- 850 instructions per step, executed once per step -- was bound by
// instruction fetch instead (2 260 cycles per step).  Threads form a kTX x kTX grid, thread (ty, tx) handles the
// elements (ty + kTX a, tx + kTX c).
#ifdef QTX_HOST_EMULATION
constexpr int kTX = 8;
#else
constexpr int kTX = 16;
#endif
static_assert(kTX * kTX == kThreads && kNB % kTX == 0, "thread grid of the shared-memory sweeps");
constexpr int kEP = kNB / kTX;  // elements per thread and dimension

// LDL^T of the leading nb x nb block of the shared tile A [kNB][kPad] (lower triangle; the upper triangle is never
// read): on exit L below the diagonal, D on it, 1 / D in dinv_s.  wv, lv: shared [kNB].  Two barriers per column.
__device__ __forceinline__ void tile_ldlt(cplx* A, int nb, cplx* wv, cplx* lv, cplx* dinv_s, int32_t* info,
                                          int64_t global_col0) {
  const int tid = threadIdx.x, tx = tid % kTX, ty = tid / kTX;
  for (int k = 0; k < nb; ++k) {
    bool zero = false;
    if (tid < nb) {  // column k: w_i = A[i][k] (what the trailing update needs), l_i = w_i / d
      cplx d = A[k * kPad + k];
      zero = (d.x == 0.0 && d.y == 0.0);
      if (zero) d = cmake(1.0, 0.0);  // reported; the factorisation continues with finite numbers
      const cplx inv = crecip(d);
      const cplx w = A[tid * kPad + k];
      const cplx l = cmul_(w, inv);
      wv[tid] = w;
      lv[tid] = l;
      if (tid > k) A[tid * kPad + k] = l;
      if (tid == k) dinv_s[k] = inv;
    }
    __syncthreads();
    if (tid == k && zero) {  // after the barrier: every thread of the column step has read the pivot
      A[k * kPad + k] = cmake(1.0, 0.0);
      if (info[0] == 0) info[0] = (int32_t)(global_col0 + k + 1);
    }
    // trailing update of rows / columns k+1 .. nb-1 (lower triangle).  All loads first, then the arithmetic, then the
    // stores: element by element the compiler has to order every store before the next load (same pointer type)
    cplx wj[kEP], li[kEP], v[kEP][kEP];
    bool m[kEP][kEP];
#pragma unroll
    for (int c = 0; c < kEP; ++c) wj[c] = wv[tx + kTX * c];
#pragma unroll
    for (int a = 0; a < kEP; ++a) li[a] = lv[ty + kTX * a];
#pragma unroll
    for (int a = 0; a < kEP; ++a)
#pragma unroll
      for (int c = 0; c < kEP; ++c) {
        const int i = ty + kTX * a, j = tx + kTX * c;
        m[a][c] = (j > k) & (j <= i) & (i < nb);
        v[a][c] = m[a][c] ? A[i * kPad + j] : cmake(0.0, 0.0);
      }
#pragma unroll
    for (int a = 0; a < kEP; ++a)
#pragma unroll
      for (int c = 0; c < kEP; ++c) {
        cfms(v[a][c], li[a], wj[c]);
        if (m[a][c]) A[(ty + kTX * a) * kPad + tx + kTX * c] = v[a][c];
      }
    __syncthreads();
  }
}

// panel solve on the shared tile X [kNB][kPad] (nr rows of A21, nb columns): on exit W = A21 L11^-T.  L11 (published
// by the diagonal CTA of this launch, read through L2) is staged kLC columns at a time in Lc [kNB][kLC + 1]; one
// barrier per column plus two per chunk.
constexpr int kLC = 16;
__device__ __forceinline__ void tile_panel(cplx* X, int nr, int nb, const cplx* L11g, int64_t ldg, cplx* Lc) {
  const int tid = threadIdx.x, tx = tid % kTX, ty = tid / kTX;
  for (int c = 0; c + 1 < nb; ++c) {  // column c is final: remove it from the columns behind it
    if (c % kLC == 0) {
      __syncthreads();  // the previous chunk is no longer read
      for (int idx = tid; idx < kNB * kLC; idx += kThreads) {
        const int j = idx / kLC, cc = idx % kLC;
        if (j < nb && c + cc < j) Lc[j * (kLC + 1) + cc] = ld_cg(L11g + (int64_t)j * ldg + c + cc);
      }
      __syncthreads();
    }
    cplx xr[kEP], l[kEP], v[kEP][kEP];
    bool m[kEP][kEP];
#pragma unroll
    for (int a = 0; a < kEP; ++a) xr[a] = X[(ty + kTX * a) * kPad + c];
#pragma unroll
    for (int q = 0; q < kEP; ++q) l[q] = Lc[(tx + kTX * q) * (kLC + 1) + c % kLC];  // rows j <= c hold stale data: masked
#pragma unroll
    for (int a = 0; a < kEP; ++a)
#pragma unroll
      for (int q = 0; q < kEP; ++q) {
        const int r = ty + kTX * a, j = tx + kTX * q;
        m[a][q] = (j > c) & (j < nb) & (r < nr);
        v[a][q] = m[a][q] ? X[r * kPad + j] : cmake(0.0, 0.0);
      }
#pragma unroll
    for (int a = 0; a < kEP; ++a)
#pragma unroll
      for (int q = 0; q < kEP; ++q) {
        cfms(v[a][q], xr[a], l[q]);
        if (m[a][q]) X[(ty + kTX * a) * kPad + tx + kTX * q] = v[a][q];
      }
    __syncthreads();
  }
}

struct StepArgs {
  cplx* M;
  int64_t n;
  int64_t k0;        // block column whose panel is applied (ignored by the head launch)
  int64_t t0;        // first row / column of the trailing matrix = block column that is factorised by this launch
  const cplx* Wprev; // [n][kNB] row-major: W of block column k0
  cplx* Wnext;       // [n][kNB]: W of block column t0
  cplx* dinvg;       // [n]: 1 / D
  unsigned* sync;    // [0] ticket, [1] flag "diagonal block t0 factorised"
  int32_t* info;
};

// kHead: no trailing update (first block column); grid = number of row tiles.  Otherwise grid = tiles of the lower
// triangle of the trailing matrix.
template <bool kHead>
__global__ void __launch_bounds__(kThreads, 2) zldlt_step_kernel(const StepArgs a) {
  QTX_DYN_SMEM(double, smd);
  __shared__ unsigned ticket_s;
  const int tid = threadIdx.x;
  if (tid == 0) ticket_s = atomicAdd(a.sync, 1u);
  __syncthreads();
  const int64_t n = a.n;
  const int T = (int)((n - a.t0 + kNB - 1) / kNB);  // row tiles of the trailing matrix
  int I, J;
  {
    const unsigned t = ticket_s;
    if (kHead || t < (unsigned)T) {
      I = (int)t;
      J = 0;
    } else {  // the remaining triangle (I, J >= 1), row by row
      const unsigned b = t - (unsigned)T;
      int r = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
      while ((unsigned)r * (unsigned)(r + 1) / 2u > b) --r;
      while ((unsigned)(r + 1) * (unsigned)(r + 2) / 2u <= b) ++r;
      I = r + 1;
      J = (int)(b - (unsigned)r * (unsigned)(r + 1) / 2u) + 1;
    }
  }
  const int64_t i0 = a.t0 + (int64_t)I * kNB, j0 = a.t0 + (int64_t)J * kNB;
  const int nr = (int)((n - i0) < kNB ? (n - i0) : kNB), nc = (int)((n - j0) < kNB ? (n - j0) : kNB);
  const TilePos tp = tile_pos();
  double tr[kFI][kFJ][2], ti[kFI][kFJ][2];
#pragma unroll
  for (int fi = 0; fi < kFI; ++fi)
#pragma unroll
    for (int fj = 0; fj < kFJ; ++fj) tr[fi][fj][0] = tr[fi][fj][1] = ti[fi][fj][0] = ti[fi][fj][1] = 0.0;

  if (!kHead) {
    // ---- acc = W L21^T over the kNB columns of block column k0: cp.async double-buffered k stages ----------------
    cplx* stage = reinterpret_cast<cplx*>(smd);  // [2 buffers][W | L][kNB][kKP]
    const int w = tid >> 5, l = tid & 31, g = l >> 2, q = l & 3;
    const int arow = (w / kWC) * kWTR + g, brow = (w % kWC) * kWTC + g;
    auto issue = [&](int chunk, int buf) {
      cplx* sW = stage + (size_t)buf * 2 * kNB * kKP;
      cplx* sL = sW + kNB * kKP;
      for (int idx = tid; idx < kNB * kKC; idx += kThreads) {
        const int r = idx / kKC, k = idx % kKC;
        const bool vw = r < nr, vl = r < nc;
        cp_async16(sW + r * kKP + k, a.Wprev + (vw ? (i0 + r) * kNB + chunk * kKC + k : 0), vw);
        cp_async16(sL + r * kKP + k, a.M + (vl ? (j0 + r) * n + a.k0 + chunk * kKC + k : 0), vl);
      }
      cp_async_commit();
    };
    constexpr int kChunks = kNB / kKC;
    issue(0, 0);
    for (int c = 0; c < kChunks; ++c) {
      if (c + 1 < kChunks) {
        issue(c + 1, (c + 1) & 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const cplx* sW = stage + (size_t)(c & 1) * 2 * kNB * kKP;
      const cplx* sL = sW + kNB * kKP;
#pragma unroll
      for (int kk = 0; kk < kKC; kk += 4) {
        cplx av[kFI], bv[kFJ];
#pragma unroll
        for (int fi = 0; fi < kFI; ++fi) av[fi] = sW[(arow + 8 * fi) * kKP + kk + q];
#pragma unroll
        for (int fj = 0; fj < kFJ; ++fj) bv[fj] = sL[(brow + 8 * fj) * kKP + kk + q];
        // four real products per complex one; consecutive MMAs go to different accumulators
#pragma unroll
        for (int fi = 0; fi < kFI; ++fi)
#pragma unroll
          for (int fj = 0; fj < kFJ; ++fj) dmma(tr[fi][fj][0], tr[fi][fj][1], av[fi].x, bv[fj].x);
#pragma unroll
        for (int fi = 0; fi < kFI; ++fi)
#pragma unroll
          for (int fj = 0; fj < kFJ; ++fj) dmma(ti[fi][fj][0], ti[fi][fj][1], av[fi].x, bv[fj].y);
#pragma unroll
        for (int fi = 0; fi < kFI; ++fi)
#pragma unroll
          for (int fj = 0; fj < kFJ; ++fj) dmma(tr[fi][fj][0], tr[fi][fj][1], -av[fi].y, bv[fj].y);
#pragma unroll
        for (int fi = 0; fi < kFI; ++fi)
#pragma unroll
          for (int fj = 0; fj < kFJ; ++fj) dmma(ti[fi][fj][0], ti[fi][fj][1], av[fi].y, bv[fj].x);
      }
      __syncthreads();  // the buffer is refilled two stages later
    }
  }
  // ---- tile = C - acc (only entries of the lower triangle of M exist) -------------------------------------------
#pragma unroll
  for (int fi = 0; fi < kFI; ++fi) {
    const int r = tp.row(fi);
#pragma unroll
    for (int fj = 0; fj < kFJ; ++fj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = tp.col(fj, e);
        cplx v = cmake(0.0, 0.0);
        if (r < nr && c < nc && j0 + c <= i0 + r) v = a.M[(i0 + r) * n + j0 + c];
        tr[fi][fj][e] = v.x - tr[fi][fj][e];
        ti[fi][fj][e] = v.y - ti[fi][fj][e];
      }
  }
  cplx* smc = reinterpret_cast<cplx*>(smd);
  if (J != 0) {  // plain trailing tile: store and leave
#pragma unroll
    for (int fi = 0; fi < kFI; ++fi) {
      const int r = tp.row(fi);
#pragma unroll
      for (int fj = 0; fj < kFJ; ++fj)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = tp.col(fj, e);
          if (r < nr && c < nc && j0 + c <= i0 + r) a.M[(i0 + r) * n + j0 + c] = cmake(tr[fi][fj][e], ti[fi][fj][e]);
        }
    }
    return;
  }
  // ---- column tile: move the updated tile from the accumulator registers into shared memory -----------------------
  cplx* Xs = smc;  // [kNB][kPad]   (the operand stages are no longer needed: every thread is past its last barrier)
#pragma unroll
  for (int fi = 0; fi < kFI; ++fi)
#pragma unroll
    for (int fj = 0; fj < kFJ; ++fj)
#pragma unroll
      for (int e = 0; e < 2; ++e) Xs[tp.row(fi) * kPad + tp.col(fj, e)] = cmake(tr[fi][fj][e], ti[fi][fj][e]);
  __syncthreads();
  const int nb = nc;  // width of block column t0
  if (I == 0) {
    // ---- the next diagonal block: factorise, publish L11 / D / 1/D ----------------------------------------------
    cplx* wv = Xs + kNB * kPad;   // [kNB]
    cplx* lv = wv + kNB;          // [kNB]
    cplx* dinv_s = lv + kNB;      // [kNB]
    tile_ldlt(Xs, nb, wv, lv, dinv_s, a.info, a.t0);
    for (int idx = tid; idx < nb * kNB; idx += kThreads) {
      const int r = idx / kNB, c = idx % kNB;
      if (c <= r) a.M[(a.t0 + r) * n + a.t0 + c] = Xs[r * kPad + c];
    }
    if (tid < nb) a.dinvg[a.t0 + tid] = dinv_s[tid];
    __syncthreads();
    if (tid == 0) publish(a.sync + 1);
    return;
  }
  // ---- a tile below it: wait for the diagonal block, then W = A21 L11^-T, L21 = W D^-1 ---------------------------
  cplx* Lc = Xs + kNB * kPad;            // [kNB][kLC + 1]
  cplx* dinv_s = Lc + kNB * (kLC + 1);   // [kNB]
  if (tid == 0) spin_until_set(a.sync + 1);
  __syncthreads();
  if (tid < nb) dinv_s[tid] = ld_cg(a.dinvg + a.t0 + tid);
  tile_panel(Xs, nr, nb, a.M + a.t0 * n + a.t0, n, Lc);
  __syncthreads();
  for (int idx = tid; idx < nr * kNB; idx += kThreads) {
    const int r = idx / kNB, c = idx % kNB;
    if (c < nb) {
      const cplx wv = Xs[r * kPad + c];
      a.Wnext[(i0 + r) * kNB + c] = wv;
      a.M[(i0 + r) * n + a.t0 + c] = cmul_(wv, dinv_s[c]);
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
  // the coefficients of block (b, j) do not depend on x_j: they are fetched into registers BEFORE waiting for its
  // flag, so that only the flag, the 64 values of x_j and 16 complex FMAs per thread sit on the dependency chain
  constexpr int kPer = kNB / kTPR;  // coefficients per thread and block
  cplx coef[kPer];
  auto fetch = [&](int j) {
    const int64_t j0 = (int64_t)j * kNB;
    const int nj = (int)((n - j0) < kNB ? (n - j0) : kNB);
#pragma unroll
    for (int t = 0; t < kPer; ++t) {
      cplx v = cmake(0.0, 0.0);
      if (!kBackward) {
        const int c = p + kTPR * t;
        if (line < nr && c < nj) v = M[(i0 + line) * n + j0 + c];
      } else {
        const int r = q + kTPR * t;
        if (col < nr && r < nj) v = M[(j0 + r) * n + i0 + col];
      }
      coef[t] = v;
    }
  };
  if (jbeg != jend) fetch(jbeg);
  for (int j = jbeg; j != jend; j += jstep) {
    if (tid == 0) spin_until_set(sync + 1 + j);  // bounded: a dependency that never arrives traps instead of hanging
    __syncthreads();
    const int64_t j0 = (int64_t)j * kNB;
    const int nj = (int)((n - j0) < kNB ? (n - j0) : kNB);
    if (tid < kNB) xj[tid] = tid < nj ? ld_cg(x + j0 + tid) : cmake(0.0, 0.0);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kPer; ++t) cfms(mine, coef[t], xj[(kBackward ? q : p) + kTPR * t]);
    if (j + jstep != jend) fetch(j + jstep);
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
  __syncthreads();
  if (tid == 0) publish(sync + 1 + b);
}

__global__ void zero_sync_kernel(unsigned* sync, int count) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) sync[i] = 0u;
}

// ---- host side --------------------------------------------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

constexpr size_t kSmemPlanes = (size_t)2 * 2 * kNB * kKP * sizeof(cplx);  // two stages of (W, L)
constexpr size_t kSmemPanelStage = (size_t)(kNB * kPad + kNB * (kLC + 1) + 3 * kNB) * sizeof(cplx);  // tile + L11 chunk
constexpr size_t kSmemStep = kSmemPlanes > kSmemPanelStage ? kSmemPlanes : kSmemPanelStage;
constexpr size_t kSmemInv = (size_t)(2 * kNB * kPad) * sizeof(cplx);
constexpr size_t kSmemTrsv = (size_t)(kNB * kPad + 2 * kNB + kNB * kTPR) * sizeof(cplx);

struct ZldltScratch {
  cplx* Wp[2];     // [n][kNB] row-major W = L21 D of the current / next block column
  cplx *dinvg, *invL;
  unsigned* fsync; // [nblk][2]: ticket and flag of every factorisation launch
  unsigned* tsync; // [2][nblk + 1]: forward and backward sweep of a solve
};

static size_t nblk_of(int64_t n) { return (size_t)((n + kNB - 1) / kNB); }
static size_t panel_bytes(int64_t n) { return align256((size_t)kNB * (size_t)n * sizeof(cplx)); }
static size_t dinv_bytes(int64_t n) { return align256((size_t)n * sizeof(cplx)); }
static size_t invl_bytes(int64_t n) { return align256(nblk_of(n) * kNB * kNB * sizeof(cplx)); }
static size_t fsync_bytes(int64_t n) { return align256(2 * nblk_of(n) * sizeof(unsigned)); }
static size_t tsync_bytes(int64_t n) { return align256(2 * (nblk_of(n) + 1) * sizeof(unsigned)); }

size_t zldlt_scratch_bytes(int64_t n) {
  return 2 * panel_bytes(n) + dinv_bytes(n) + invl_bytes(n) + fsync_bytes(n) + tsync_bytes(n) + 256;
}

static ZldltScratch carve(void* scratch, int64_t n) {
  char* p = (char*)align256((size_t)scratch);
  ZldltScratch s;
  s.Wp[0] = (cplx*)p;
  p += panel_bytes(n);
  s.Wp[1] = (cplx*)p;
  p += panel_bytes(n);
  s.dinvg = (cplx*)p;
  p += dinv_bytes(n);
  s.invL = (cplx*)p;
  p += invl_bytes(n);
  s.fsync = (unsigned*)p;
  p += fsync_bytes(n);
  s.tsync = (unsigned*)p;
  return s;
}

static int zldlt_prepare() {
#ifndef QTX_HOST_EMULATION
  static thread_local int prepared_device = -1;  // the attribute is per function and device
  int dev = 0;
  QTX_CUDA(cudaGetDevice(&dev));
  if (prepared_device != dev) {
    QTX_CUDA(cudaFuncSetAttribute(zldlt_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemStep));
    QTX_CUDA(cudaFuncSetAttribute(zldlt_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemStep));
    QTX_CUDA(cudaFuncSetAttribute(zldlt_diaginv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemInv));
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
  const unsigned nblk = (unsigned)nblk_of(n);
  QTX_LAUNCH_SMEM(zero_sync_kernel, 1, 256, 0, st, s.fsync, (int)(2 * nblk));
  QTX_LAUNCH_CHECK();
  StepArgs a;
  a.M = M;
  a.n = n;
  a.dinvg = s.dinvg;
  a.info = info;
  // block column 0: factorise the first diagonal block and its panel
  a.k0 = 0;
  a.t0 = 0;
  a.Wprev = nullptr;
  a.Wnext = s.Wp[0];
  a.sync = s.fsync;
  QTX_LAUNCH_SMEM(zldlt_step_kernel<true>, nblk, kThreads, kSmemStep, st, a);
  QTX_LAUNCH_CHECK();
  for (unsigned step = 1; step < nblk; ++step) {  // apply block column step - 1, factorise block column step
    a.k0 = (int64_t)(step - 1) * kNB;
    a.t0 = (int64_t)step * kNB;
    a.Wprev = s.Wp[(step - 1) & 1];
    a.Wnext = s.Wp[step & 1];
    a.sync = s.fsync + 2 * step;
    const unsigned T = nblk - step;
    QTX_LAUNCH_SMEM(zldlt_step_kernel<false>, T * (T + 1) / 2, kThreads, kSmemStep, st, a);
    QTX_LAUNCH_CHECK();
  }
  QTX_LAUNCH_SMEM(zldlt_diaginv_kernel, nblk, kThreads, kSmemInv, st, M, n, s.invL);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

int zldlt_solve(const cuDoubleComplex* M, int64_t n, cuDoubleComplex* x, void* scratch, cudaStream_t st) {
  int rc = zldlt_prepare();
  if (rc) return rc;
  const ZldltScratch s = carve(scratch, n);
  const unsigned nblk = (unsigned)nblk_of(n);
  QTX_LAUNCH_SMEM(zero_sync_kernel, 1, 256, 0, st, s.tsync, (int)(2 * (nblk + 1)));
  QTX_LAUNCH_CHECK();
  QTX_LAUNCH_SMEM(ztrsv_kernel<false>, nblk, kThreads, kSmemTrsv, st, M, s.invL, n, x, s.tsync);
  QTX_LAUNCH_CHECK();
  QTX_LAUNCH_SMEM(ztrsv_kernel<true>, nblk, kThreads, kSmemTrsv, st, M, s.invL, n, x, s.tsync + nblk + 1);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

int zldlt_block_size() { return kNB; }

}  // namespace qtx
