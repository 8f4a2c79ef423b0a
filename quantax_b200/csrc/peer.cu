// Peer-memory plumbing of the fused Gram + exchange path (distributed MinSR, quantax/optimizer/solver.py:134-139:
// `T = Adag^T Adag` on column-sharded operands is a local Gram followed by a sum over devices).
//
// Every rank owns a staging area  stage[P][ns][ns]  (float64, only j <= i is used) and a flag array
// flags[P] (uint64 epochs), both plain cudaMalloc memory exported with CUDA IPC and mapped by the other
// ranks of the node.  The Gram kernel of rank r (gram_tc2.cu, PUSH variant) stores every finished tile of
// its partial T straight into stage_q[r] of every peer q over NVLink while the tensor cores work on the
// next tile; when the kernel is done a one-thread kernel publishes flags_q[r] = epoch with release
// semantics at system scope.  gram_reduce_kernel then waits (acquire, bounded) until all P flags of its own
// rank carry the epoch and sums the P partials in RANK ORDER, so that all ranks hold bit-identical T
// (the replicated eigh must see the same matrix everywhere).  No NCCL call is involved in the data path.
//
// Re-use of the staging area across solves is ordered by the collectives that surround the Gram in
// distributed_minnorm (the all-gather of x after it and the all-to-all of Obar before the next one): a rank
// cannot start pushing step k+1 before every peer has entered that all-to-all, i.e. finished reducing step k.
#include "common.cuh"

namespace qtx {

__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct PeerPtrs {
  uint64_t* flags[QTX_MAX_PEERS];
};

// flags_q[rank] = epoch on every rank q (own included); runs after the pushing kernel in stream order
__global__ void peer_signal_kernel(PeerPtrs peers, int nranks, int rank, uint64_t epoch) {
  if (threadIdx.x < nranks) {
    __threadfence_system();
    st_release_sys_u64(peers.flags[threadIdx.x] + rank, epoch);
  }
}

struct ReduceSrc {
  const double* src[QTX_MAX_PEERS];  // src[q] = partial of rank q as seen by this rank ([ns, ns], j <= i valid)
};

// T[i, j] = T[j, i] = sum_q src[q][i, j] for j <= i, in rank order.  Persistent CTAs walk the 32 x 32 blocks of the
// lower triangle (t enumerates bi >= bj); the mirrored block is written through shared memory so that both stores
// are coalesced.  Thread 0 first waits until every rank has published `epoch`; the wait is bounded (trap
// instead of a hang if a peer died).
__global__ void __launch_bounds__(256) gram_reduce_kernel(ReduceSrc srcs, int nranks, int64_t ns, double* T,
                                                          const uint64_t* __restrict__ flags, uint64_t epoch,
                                                          uint64_t timeout_ns, int64_t nblocks) {
  __shared__ double tile[32][33];
  if (threadIdx.x == 0 && flags) {
    const uint64_t t0 = global_timer_ns();
    for (int q = 0; q < nranks; ++q) {
      while (ld_acquire_sys_u64(flags + q) < epoch) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > timeout_ns) __trap();
      }
    }
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int64_t t = blockIdx.x; t < nblocks; t += gridDim.x) {
    // block index -> (bi, bj), bj <= bi
    int64_t bi = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while (bi * (bi + 1) / 2 > t) --bi;
    const int64_t bj = t - bi * (bi + 1) / 2;
    const bool diag = bi == bj;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < nranks; ++q) {  // rank order; the four rows of a thread are independent loads
      const double* src = srcs.src[q];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t i = bi * 32 + ty + 8 * r, j = bj * 32 + tx;
        if (i < ns && j <= i) acc[r] += __ldcg(src + i * ns + j);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int li = ty + 8 * r;
      const int64_t i = bi * 32 + li, j = bj * 32 + tx;
      if (i < ns && j <= i) T[i * ns + j] = acc[r];
      tile[li][tx] = acc[r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int lj = ty + 8 * r;  // row of the mirrored block = column of the source block
      const int64_t jj = bj * 32 + lj, ii = bi * 32 + tx;
      // mirrored element T[jj, ii] = value(ii, jj); on a diagonal block only the strict upper part is missing
      if (ii < ns && jj < ns && (diag ? jj < ii : true)) T[jj * ns + ii] = tile[tx][lj];
    }
    __syncthreads();  // the tile is reused by the next block of this CTA
  }
}

}  // namespace qtx

using namespace qtx;

// ---- CUDA IPC helpers ------------------------------------------------------------------------------------
extern "C" int qtx_peer_alloc(size_t bytes, void** ptr_out) {
  QTX_REQUIRE(ptr_out && bytes > 0, QTX_ERR_INVALID, "qtx_peer_alloc: bad argument");
  QTX_CUDA(cudaMalloc(ptr_out, bytes));
  QTX_CUDA(cudaMemset(*ptr_out, 0, bytes));
  QTX_CUDA(cudaDeviceSynchronize());
  return QTX_OK;
}

extern "C" int qtx_peer_free(void* ptr) {
  if (ptr) QTX_CUDA(cudaFree(ptr));
  return QTX_OK;
}

extern "C" int qtx_peer_export(void* ptr, void* handle64_out) {
  QTX_REQUIRE(ptr && handle64_out, QTX_ERR_INVALID, "qtx_peer_export: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  QTX_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64_out, &h, sizeof(h));
  return QTX_OK;
}

extern "C" int qtx_peer_open(const void* handle64, void** ptr_out) {
  QTX_REQUIRE(handle64 && ptr_out, QTX_ERR_INVALID, "qtx_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  QTX_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return QTX_OK;
}

extern "C" int qtx_peer_close(void* ptr) {
  if (ptr) QTX_CUDA(cudaIpcCloseMemHandle(ptr));
  return QTX_OK;
}

// ---- exchange steps ----------------------------------------------------------------------------------------
extern "C" int qtx_peer_signal(void* const* peer_flags, int nranks, int rank, uint64_t epoch, qtx_stream_t stream) {
  QTX_REQUIRE(peer_flags && nranks >= 1 && nranks <= QTX_MAX_PEERS && rank >= 0 && rank < nranks, QTX_ERR_INVALID,
              "qtx_peer_signal: bad argument");
  PeerPtrs pp{};
  for (int q = 0; q < nranks; ++q) {
    QTX_REQUIRE(peer_flags[q], QTX_ERR_INVALID, "qtx_peer_signal: null flag pointer for rank %d", q);
    pp.flags[q] = (uint64_t*)peer_flags[q];
  }
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, nranks, rank, epoch);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_gram_reduce(const void* const* partials, int nranks, int64_t ns, double* T_out,
                               const void* my_flags, uint64_t epoch, double timeout_s, qtx_stream_t stream) {
  QTX_REQUIRE(partials && T_out && nranks >= 1 && nranks <= QTX_MAX_PEERS && ns > 0, QTX_ERR_INVALID,
              "qtx_gram_reduce: bad argument");
  ReduceSrc rs{};
  for (int q = 0; q < nranks; ++q) {
    QTX_REQUIRE(partials[q], QTX_ERR_INVALID, "qtx_gram_reduce: null partial for rank %d", q);
    rs.src[q] = (const double*)partials[q];
  }
  const int64_t nb = (ns + 31) / 32;
  const int64_t nblocks = nb * (nb + 1) / 2;
  QTX_REQUIRE(nblocks < (int64_t)1 << 31, QTX_ERR_UNSUPPORTED, "qtx_gram_reduce: ns too large");
  if (timeout_s <= 0) timeout_s = 60.0;
  const int64_t grid = nblocks < 8ll * num_sms() ? nblocks : 8ll * num_sms();  // persistent CTAs: one flag wait each
  gram_reduce_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(rs, nranks, ns, T_out,
                                                                       (const uint64_t*)my_flags, epoch,
                                                                       (uint64_t)(timeout_s * 1e9), nblocks);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
