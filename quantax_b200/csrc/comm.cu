// Collectives of the data-parallel VMC step behind the C ABI (qtx_comm_*), so that any host -- the torch binding of
// this repo, a jax.ffi binder -- drives the multi-GPU MinSR solve without its own communication layer, and the
// distributed solve itself (qtx_minsr_solve_dist): the row-sharded -> column-sharded exchange of Obar, the sum of the
// partial Gram matrices, the rank split of the shifted solves and the gathers of y and x that the reference gets from
// GSPMD (quantax/optimizer/solver.py:131-147).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process -- torch's -- or the system one): the
// library loads, and every single-GPU entry point works, on a machine without NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace qtx {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_api(NcclApi** out) {
  NcclApi& a = g_nccl;
  if (!a.handle) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host framework already loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    QTX_REQUIRE(h, QTX_ERR_UNSUPPORTED, "qtx_comm: libnccl.so.2 not found (%s)", dlerror());
#define QTX_NCCL_SYM(field, name)                                                             \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                              \
  QTX_REQUIRE(a.field, QTX_ERR_UNSUPPORTED, "qtx_comm: symbol %s missing from libnccl", name)
    QTX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    QTX_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    QTX_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    QTX_NCCL_SYM(CommCount, "ncclCommCount");
    QTX_NCCL_SYM(CommUserRank, "ncclCommUserRank");
    QTX_NCCL_SYM(AllReduce, "ncclAllReduce");
    QTX_NCCL_SYM(AllGather, "ncclAllGather");
    QTX_NCCL_SYM(Broadcast, "ncclBroadcast");
    QTX_NCCL_SYM(Send, "ncclSend");
    QTX_NCCL_SYM(Recv, "ncclRecv");
    QTX_NCCL_SYM(GroupStart, "ncclGroupStart");
    QTX_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    QTX_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef QTX_NCCL_SYM
    a.handle = h;
  }
  *out = &a;
  return QTX_OK;
}

#define QTX_NCCL(api, expr)                                                                       \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess) {                                                                      \
      qtx::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
      return QTX_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

struct Comm {
  ncclComm_t nccl;
  int nranks, rank;
  bool owned;
};

// send [P][nl][npc] <- A [nl, ld] with columns [p npc, (p+1) npc) of the zero-padded parameter axis
template <typename T>
__global__ void __launch_bounds__(256) pack_column_shards_kernel(const T* __restrict__ A, int64_t nl, int64_t np,
                                                                 int64_t ld, int P, int64_t npc, T* __restrict__ send) {
  const int64_t total = (int64_t)P * nl * npc;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx % npc, r = (idx / npc) % nl, p = idx / (npc * nl);
    const int64_t col = p * npc + c;
    send[idx] = col < np ? A[r * ld + col] : T(0);
  }
}

__global__ void abs_info_kernel(int32_t* info) { info[0] = info[0] < 0 ? -info[0] : info[0]; }

// optional phase timing of qtx_minsr_solve_dist (qtx_minsr_solve_dist_timing): CUDA events on the solve's stream at the
// phase boundaries; nothing is recorded unless enabled
constexpr int kDistPhases = 8;
static bool g_dist_timing = false;
static bool g_dist_timed = false;
static cudaEvent_t g_dist_ev[kDistPhases + 1];
static bool g_dist_ev_ready = false;
static void dist_mark(int i, cudaStream_t st) {
  if (!g_dist_timing) return;
  if (!g_dist_ev_ready) {
    for (int k = 0; k <= kDistPhases; ++k) cudaEventCreate(&g_dist_ev[k]);
    g_dist_ev_ready = true;
  }
  cudaEventRecord(g_dist_ev[i], st);
  if (i == kDistPhases) g_dist_timed = true;
}

}  // namespace qtx

using namespace qtx;

extern "C" int qtx_comm_unique_id(void* id_out_128_bytes) {
  QTX_REQUIRE(id_out_128_bytes, QTX_ERR_INVALID, "qtx_comm_unique_id: null output");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  QTX_NCCL(api, api->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id_out_128_bytes)));
  return QTX_OK;
}

extern "C" int qtx_comm_init(qtx_comm_t* comm_out, int nranks, int rank, const void* id_128_bytes) {
  QTX_REQUIRE(comm_out && id_128_bytes && nranks >= 1 && rank >= 0 && rank < nranks, QTX_ERR_INVALID,
              "qtx_comm_init: bad argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id_128_bytes, sizeof(id));
  ncclComm_t c;
  QTX_NCCL(api, api->CommInitRank(&c, nranks, id, rank));
  *comm_out = new Comm{c, nranks, rank, true};
  return QTX_OK;
}

// wrap a communicator the host already has (an ncclComm_t handed over by the framework); not destroyed by qtx_comm_destroy
extern "C" int qtx_comm_adopt(qtx_comm_t* comm_out, void* nccl_comm) {
  QTX_REQUIRE(comm_out && nccl_comm, QTX_ERR_INVALID, "qtx_comm_adopt: bad argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  int n = 0, r = 0;
  QTX_NCCL(api, api->CommCount((ncclComm_t)nccl_comm, &n));
  QTX_NCCL(api, api->CommUserRank((ncclComm_t)nccl_comm, &r));
  *comm_out = new Comm{(ncclComm_t)nccl_comm, n, r, false};
  return QTX_OK;
}

extern "C" int qtx_comm_destroy(qtx_comm_t comm) {
  if (!comm) return QTX_OK;
  Comm* c = (Comm*)comm;
  if (c->owned) {
    NcclApi* api;
    int rc = nccl_api(&api);
    if (rc) return rc;
    QTX_NCCL(api, api->CommDestroy(c->nccl));
  }
  delete c;
  return QTX_OK;
}

extern "C" int qtx_comm_size(qtx_comm_t comm) { return comm ? ((Comm*)comm)->nranks : 0; }
extern "C" int qtx_comm_rank(qtx_comm_t comm) { return comm ? ((Comm*)comm)->rank : -1; }

static int nccl_type(int dtype, ncclDataType_t* t) {
  switch (dtype) {
    case QTX_F32: *t = ncclFloat32; return QTX_OK;
    case QTX_F64: *t = ncclFloat64; return QTX_OK;
    case QTX_I32: *t = ncclInt32; return QTX_OK;
    default: set_error("qtx_comm: unsupported dtype %d", dtype); return QTX_ERR_INVALID;
  }
}

extern "C" int qtx_comm_all_reduce(qtx_comm_t comm, const void* send, void* recv, int64_t count, int dtype, int op,
                                   qtx_stream_t stream) {
  QTX_REQUIRE(comm && send && recv && count >= 0 && (op == QTX_SUM || op == QTX_MAX), QTX_ERR_INVALID,
              "qtx_comm_all_reduce: bad argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  ncclDataType_t t;
  rc = nccl_type(dtype, &t);
  if (rc) return rc;
  QTX_NCCL(api, api->AllReduce(send, recv, (size_t)count, t, op == QTX_SUM ? ncclSum : ncclMax, ((Comm*)comm)->nccl,
                               (cudaStream_t)stream));
  count_launch();
  return QTX_OK;
}

extern "C" int qtx_comm_all_gather(qtx_comm_t comm, const void* send, void* recv, int64_t bytes_per_rank,
                                   qtx_stream_t stream) {
  QTX_REQUIRE(comm && send && recv && bytes_per_rank >= 0, QTX_ERR_INVALID, "qtx_comm_all_gather: bad argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  QTX_NCCL(api, api->AllGather(send, recv, (size_t)bytes_per_rank, ncclUint8, ((Comm*)comm)->nccl, (cudaStream_t)stream));
  count_launch();
  return QTX_OK;
}

extern "C" int qtx_comm_broadcast(qtx_comm_t comm, void* buf, int64_t bytes, int root, qtx_stream_t stream) {
  QTX_REQUIRE(comm && buf && bytes >= 0 && root >= 0 && root < ((Comm*)comm)->nranks, QTX_ERR_INVALID,
              "qtx_comm_broadcast: bad argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  QTX_NCCL(api, api->Broadcast(buf, buf, (size_t)bytes, ncclUint8, root, ((Comm*)comm)->nccl, (cudaStream_t)stream));
  count_launch();
  return QTX_OK;
}

// recv block p (bytes_per_peer bytes) <- send block `rank` of peer p
extern "C" int qtx_comm_all_to_all(qtx_comm_t comm, const void* send, void* recv, int64_t bytes_per_peer,
                                   qtx_stream_t stream) {
  QTX_REQUIRE(comm && send && recv && bytes_per_peer >= 0, QTX_ERR_INVALID, "qtx_comm_all_to_all: bad argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  Comm* c = (Comm*)comm;
  QTX_NCCL(api, api->GroupStart());
  for (int p = 0; p < c->nranks; ++p) {
    QTX_NCCL(api, api->Send((const char*)send + (size_t)p * bytes_per_peer, (size_t)bytes_per_peer, ncclUint8, p, c->nccl,
                            (cudaStream_t)stream));
    QTX_NCCL(api, api->Recv((char*)recv + (size_t)p * bytes_per_peer, (size_t)bytes_per_peer, ncclUint8, p, c->nccl,
                            (cudaStream_t)stream));
  }
  QTX_NCCL(api, api->GroupEnd());
  count_launch();
  return QTX_OK;
}

// ---- distributed MinSR solve --------------------------------------------------------------------------------------
namespace qtx {
struct DistLayout {
  size_t send, recv, T, bfull, ydd, yall, y, xc, lam, info, gram, pinv, total;
  size_t gram_bytes, pinv_bytes;
  int64_t npc;
};
static size_t al(size_t v) { return (v + 255) & ~(size_t)255; }
static int shift_mask_of(int P, int rank) {  // optimizer.rational_shift_masks
  if (P <= 1) return 7;
  if (P == 2) return rank == 0 ? 5 : 2;
  return rank < 3 ? (1 << rank) : 0;
}
static int dist_layout(int P, int rank, int dtype, int64_t nl, int64_t np, int nslices, DistLayout* L) {
  const size_t esz = dtype == QTX_F32 ? 4 : 8;
  const int64_t ns = nl * P;
  L->npc = (np + P - 1) / P;
  int mask = shift_mask_of(P, rank), nsh = 0;
  for (int k = 0; k < 3; ++k) nsh += (mask >> k) & 1;
  L->gram_bytes = qtx_gram_workspace_size(dtype, ns, L->npc, nslices);
  L->pinv_bytes = qtx_pinv_ldlt_workspace_size(ns, nsh > 0 ? nsh : 1);
  if (L->pinv_bytes == 0) return QTX_ERR_INVALID;
  size_t off = 0;
  L->send = off; off += al((size_t)P * nl * L->npc * esz);
  L->recv = off; off += al((size_t)P * nl * L->npc * esz);
  L->T = off; off += al((size_t)ns * ns * 8);
  L->bfull = off; off += al((size_t)ns * 8);
  L->ydd = off; off += al((size_t)2 * ns * 8);
  L->yall = off; off += al((size_t)P * 2 * ns * 8);
  L->y = off; off += al((size_t)ns * 8);
  L->xc = off; off += al((size_t)L->npc * 8);
  L->lam = off; off += 256;
  L->info = off; off += 256;
  L->gram = off; off += al(L->gram_bytes);
  L->pinv = off; off += al(L->pinv_bytes);
  L->total = off + 256;
  return QTX_OK;
}
}  // namespace qtx

extern "C" size_t qtx_minsr_solve_dist_workspace_size(qtx_comm_t comm, int dtype, int64_t nl, int64_t np, int nslices) {
  if (!comm || nl <= 0 || np <= 0 || (dtype != QTX_F32 && dtype != QTX_F64)) return 0;
  Comm* c = (Comm*)comm;
  DistLayout L;
  if (dist_layout(c->nranks, c->rank, dtype, nl, np, nslices, &L)) return 0;
  return L.total;
}

extern "C" int qtx_minsr_solve_dist(qtx_comm_t comm, int dtype, const void* A_local, int64_t nl, int64_t np, int64_t ld,
                                    const double* b_local, double rtol, double atol, int nslices, int lanczos_steps,
                                    int refine_steps, double* x_out, int32_t* info_out, void* workspace,
                                    size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(comm && A_local && b_local && x_out && info_out && workspace && nl > 0 && np > 0 && ld >= np &&
                  (dtype == QTX_F32 || dtype == QTX_F64) && lanczos_steps != 0,
              QTX_ERR_INVALID, "qtx_minsr_solve_dist: bad argument");
  Comm* c = (Comm*)comm;
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  const int P = c->nranks, rank = c->rank;
  const int64_t ns = nl * P;
  DistLayout L;
  rc = dist_layout(P, rank, dtype, nl, np, nslices, &L);
  QTX_REQUIRE(rc == QTX_OK && workspace_bytes >= L.total, QTX_ERR_INVALID, "qtx_minsr_solve_dist: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)al((size_t)workspace);
  const size_t esz = dtype == QTX_F32 ? 4 : 8;
  const int64_t npc = L.npc;
  dist_mark(0, st);
  // 1. row-sharded -> column-sharded (solver.py:134-137; the parameter axis is zero-padded to a multiple of P)
  const unsigned grid = 8u * (unsigned)num_sms();
  if (dtype == QTX_F32)
    pack_column_shards_kernel<float><<<grid, 256, 0, st>>>((const float*)A_local, nl, np, ld, P, npc, (float*)(base + L.send));
  else
    pack_column_shards_kernel<double><<<grid, 256, 0, st>>>((const double*)A_local, nl, np, ld, P, npc, (double*)(base + L.send));
  QTX_LAUNCH_CHECK();
  rc = qtx_comm_all_to_all(comm, base + L.send, base + L.recv, (int64_t)((size_t)nl * npc * esz), stream);
  if (rc) return rc;
  dist_mark(1, st);
  // 2. Gram of the column shard (all Ns rows, rank-major = global sample order), summed over the ranks (solver.py:139)
  double* T = (double*)(base + L.T);
  rc = qtx_gram(dtype, base + L.recv, ns, npc, npc, nslices, T, 0, base + L.gram, L.gram_bytes, stream);
  if (rc) return rc;
  dist_mark(2, st);
  QTX_NCCL(api, api->AllReduce(T, T, (size_t)ns * ns, ncclFloat64, ncclSum, c->nccl, st));
  double* bfull = (double*)(base + L.bfull);
  QTX_NCCL(api, api->AllGather(b_local, bfull, (size_t)nl, ncclFloat64, c->nccl, st));
  dist_mark(3, st);
  // 3. soft pseudo-inverse y = f(T) b: the ranks take different shifts (T, b are identical everywhere)
  const int mask = shift_mask_of(P, rank);
  int nsh = 0;
  for (int k = 0; k < 3; ++k) nsh += (mask >> k) & 1;
  double* lam = (double*)(base + L.lam);
  if (lanczos_steps < 0) {  // exactly |lanczos_steps| steps, no host read-back
    rc = qtx_sym_absmax_eig_ws(T, ns, 0, -lanczos_steps, lam, base + L.pinv, L.pinv_bytes, nsh > 0 ? nsh : 1, stream);
    if (rc) return rc;
  } else {
    // adaptive like the single-GPU solve (optimizer.py sym_absmax_eig): continue the recurrence to 32, 64, 128, ...
    // steps until two consecutive values agree to 1e-7 (the error roughly squares when the step count doubles); one
    // scalar read-back per stage.  T is bit-identical on all ranks, so all ranks stop at the same stage.
    const int cap = (int)(lanczos_steps < ns ? lanczos_steps : ns);
    double prev = 0.0;
    bool have_prev = false;
    int done = 0;
    for (int upto = 32;; upto *= 2) {
      if (upto > cap) upto = cap;
      rc = qtx_sym_absmax_eig_ws(T, ns, done, upto, lam, base + L.pinv, L.pinv_bytes, nsh > 0 ? nsh : 1, stream);
      if (rc) return rc;
      done = upto;
      if (upto >= cap) break;
      double cur = 0.0;
      QTX_CUDA(cudaMemcpyAsync(&cur, lam, sizeof(double), cudaMemcpyDeviceToHost, st));
      QTX_CUDA(cudaStreamSynchronize(st));
      if (have_prev && fabs(cur - prev) <= 1e-7 * fabs(cur)) break;
      prev = cur;
      have_prev = true;
    }
  }
  dist_mark(4, st);
  double* ydd = (double*)(base + L.ydd);
  int32_t* info = (int32_t*)(base + L.info);
  rc = qtx_pinv_ldlt_partial(T, ns, bfull, rtol, atol, lam, mask, refine_steps, ydd, 0, info, base + L.pinv, L.pinv_bytes,
                             stream);
  if (rc) return rc;
  abs_info_kernel<<<1, 1, 0, st>>>(info);
  QTX_LAUNCH_CHECK();
  dist_mark(5, st);
  double* yall = (double*)(base + L.yall);
  QTX_NCCL(api, api->AllGather(ydd, yall, (size_t)2 * ns, ncclFloat64, c->nccl, st));
  QTX_NCCL(api, api->AllReduce(info, info_out, 1, ncclInt32, ncclMax, c->nccl, st));
  double* y = (double*)(base + L.y);
  rc = qtx_dd_sum_scale(yall, P, ns, 1.0 / 3.0, y, stream);
  if (rc) return rc;
  dist_mark(6, st);
  // 4. column shard of x = A^T y, gathered (solver.py:146)
  double* xc = (double*)(base + L.xc);
  rc = qtx_matvec_t(dtype, base + L.recv, ns, npc, npc, y, xc, 0, stream);
  if (rc) return rc;
  dist_mark(7, st);
  if ((int64_t)P * npc == np) {
    QTX_NCCL(api, api->AllGather(xc, x_out, (size_t)npc, ncclFloat64, c->nccl, st));
  } else {  // gather the padded vector into the send area (free again), copy the first np entries
    double* xpad = (double*)(base + L.send);
    QTX_NCCL(api, api->AllGather(xc, xpad, (size_t)npc, ncclFloat64, c->nccl, st));
    QTX_CUDA(cudaMemcpyAsync(x_out, xpad, (size_t)np * 8, cudaMemcpyDeviceToDevice, st));
  }
  dist_mark(8, st);
  count_launch(5);
  return QTX_OK;
}

extern "C" int qtx_minsr_solve_dist_timing(int enable) {
  g_dist_timing = enable != 0;
  g_dist_timed = false;
  return QTX_OK;
}

extern "C" int qtx_minsr_solve_dist_phases(double* ms_out_8) {
  QTX_REQUIRE(ms_out_8, QTX_ERR_INVALID, "qtx_minsr_solve_dist_phases: null output");
  QTX_REQUIRE(g_dist_timing && g_dist_timed, QTX_ERR_INVALID, "qtx_minsr_solve_dist_phases: no timed solve (enable with qtx_minsr_solve_dist_timing)");
  QTX_CUDA(cudaEventSynchronize(g_dist_ev[kDistPhases]));
  for (int k = 0; k < kDistPhases; ++k) {
    float ms = 0.f;
    QTX_CUDA(cudaEventElapsedTime(&ms, g_dist_ev[k], g_dist_ev[k + 1]));
    ms_out_8[k] = (double)ms;
  }
  return QTX_OK;
}
