// 2-CTA variant of the tensor-core Gram kernel: a CTA pair (cluster 2x1x1, same TPC) computes a
// 256 x 128 output tile with tcgen05.mma.cta_group::2 (UMMA M = 256).  Each CTA stages only its own
// 128 rows of the I panel and HALF (64 rows) of the J panel, so the L2 -> shared-memory traffic per
// output drops to 0.75x of the single-CTA kernel (which is bound by exactly that traffic) and the
// shared-memory operand reads per MMA drop from 8 KB to 6 KB per CTA.
//   - both CTAs run a TMA producer; their loads signal the LEADER's full barrier
//     (cp.async.bulk.tensor ... .cta_group::2, barrier address with the peer bit cleared);
//   - only the leader issues MMAs; tcgen05.commit ... .multicast::cluster releases the smem stage and
//     publishes the accumulators in BOTH CTAs;
//   - each CTA's epilogue warps drain their own 128 TMEM lanes; the peer's warps arrive remotely on the
//     leader's tmem_empty barrier.
#include <stdlib.h>

#include "gram_tc_common.cuh"

namespace qtx {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address -> same offset in the even CTA
constexpr int kSliceBytesB2 = (kTile / 2) * kBK;  // 64 rows x 32 B

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ bool elect_one_lane() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// instruction descriptor: D = S32, A = B = signed int8, K-major, M = 256 (2 CTAs), N = 128
constexpr uint32_t kIdescI8_2sm = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTile >> 3) << 17) |
                                  ((uint32_t)(256 >> 4) << 24);

template <int S, int PASS>
__device__ __forceinline__ void issue_kblock2(uint32_t tmem_base, uint32_t desc_lo, uint32_t desc_hi, uint32_t acc_first) {
  constexpr int d_lo = 2 + PASS * kLevelsPerPass;
  constexpr int d_hi = (d_lo + kLevelsPerPass - 1 < S + 1) ? d_lo + kLevelsPerPass - 1 : S + 1;
#pragma unroll
  for (int d = d_lo; d <= d_hi; ++d) {
    const int a_lo = (d - S > 1) ? d - S : 1, a_hi = (d - 1 < S) ? d - 1 : S;
#pragma unroll
    for (int a = a_lo; a <= a_hi; ++a) {
      const int b = d - a;
      // stage layout: [A slices: S x 4 KB][B half-slices: S x 2 KB]
      const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo + (uint32_t)(a - 1) * (kSliceBytes >> 4));
      const uint64_t db = ((uint64_t)desc_hi << 32) |
                          (uint64_t)(desc_lo + (uint32_t)S * (kSliceBytes >> 4) + (uint32_t)(b - 1) * (kSliceBytesB2 >> 4));
      umma_i8_2sm(tmem_base + (uint32_t)(d - d_lo) * kTile, da, db, kIdescI8_2sm, a == a_lo ? acc_first : 1u);
    }
  }
}

// tile t of the pair -> (I2: 256-row block, J: 128-column block), J <= 2*I2 + 1.  With a tile map (opt-in, see
// gram_tc.cu) the order is the host's supertile order: the tiles that run at the same time (one per CTA pair, round
// robin) form a 6 x 12 block of the lower triangle, i.e. share 6 row panels and 12 column panels instead of 1 and 74
__device__ __forceinline__ void tile2_from_index(const int* __restrict__ map, int t, int& I2, int& J) {
  if (map) {
    const int m = __ldg(map + t);
    I2 = m >> 16;
    J = m & 0xffff;
    return;
  }
  int i = (int)((sqrtf(4.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((i + 1) * (i + 2) <= t) ++i;
  while (i * (i + 1) > t) --i;
  I2 = i;
  J = t - i * (i + 1);
}

// Peers of the fused Gram + exchange path (peer.cu): slot[q] is where rank q expects THIS rank's partial T
// ([ns, ns] float64, only j <= i is written); slot[rank] is unused (the own partial stays in p.T).
struct GramPushParams {
  int nranks;
  int rank;
  double* slot[QTX_MAX_PEERS];
};

template <int S, bool PUSH>
__device__ __forceinline__ void gram_tc2_body(const CUtensorMap& tmapA, const CUtensorMap& tmapB, const GramTcParams& p,
                                              const GramPushParams& pp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t stage_bytes = (uint32_t)S * (kSliceBytes + kSliceBytesB2);
  constexpr int npasses = (S + kLevelsPerPass - 1) / kLevelsPerPass;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  double* stg = reinterpret_cast<double*>(((uintptr_t)(tmem_holder + 1) + 15) & ~(uintptr_t)15);  // [4 epilogue warps][32][17]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full_bar + i, 1);   // leader's arrive.expect_tx (bytes of both CTAs)
      mbar_init(empty_bar + i, 1);  // multicast commit from the leader's MMA thread
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);  // 4 epilogue warps of each CTA (used in the leader only)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc2(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrival / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int nb2 = p.nb / 2;               // 256-row blocks
  const int ntiles = nb2 * (nb2 + 1);
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = cluster_id; t < ntiles; t += nclusters) {
        int I2, J;
        tile2_from_index(p.tile_map, t, I2, J);
        const int rowA = I2 * 256 + (int)rank * kTile;
        const int rowB = J * kTile + (int)rank * (kTile / 2);
#pragma unroll
        for (int pass = 0; pass < npasses; ++pass) {
          const int d_hi = min(2 + (pass + 1) * kLevelsPerPass - 1, S + 1);
          const int nsl = min(S, d_hi - 1);
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            unsigned char* base = smem + (size_t)stage * stage_bytes;
            if (leader) mbar_expect_tx(full_bar + stage, 2u * nsl * (kSliceBytes + kSliceBytesB2));
#pragma unroll
            for (int a = 0; a < S; ++a) {
              if (a < nsl) {
                tma_load_3d_2sm(base + (size_t)a * kSliceBytes, &tmapA, full_bar + stage, 0, rowA, a * p.nkb + kb);
                tma_load_3d_2sm(base + (size_t)S * kSliceBytes + (size_t)a * kSliceBytesB2, &tmapB, full_bar + stage, 0,
                                rowB, a * p.nkb + kb);
              }
            }
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp of the leader CTA runs the (warp-uniform) loop and one elected lane issues,
    // so that ptxas keeps the descriptors in uniform registers (UTCIMMA takes uniform-register operands; with a
    // lane-0 branch every MMA paid a vector->uniform election loop, ~70 cycles per MMA of 64 tensor cycles) =====
    if (leader) {
      uint32_t stage = 0, phase = 0, tphase = 0;
      const uint64_t desc0 = make_desc_sw32(smem_u32(smem));
      const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
      for (int t = cluster_id; t < ntiles; t += nclusters) {
#pragma unroll
        for (int pass = 0; pass < npasses; ++pass) {
          mbar_wait(tmem_empty, tphase ^ 1);  // both CTAs' epilogues have drained the accumulators
          tc_fence_after();
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            const uint32_t desc_lo = desc_lo0 + stage * (stage_bytes >> 4);
            const uint32_t acc_first = kb > 0 ? 1u : 0u;
            if (elect_one_lane()) {
              if (pass == 0) issue_kblock2<S, 0>(tmem_base, desc_lo, desc_hi, acc_first);
              else issue_kblock2<S, (npasses > 1 ? 1 : 0)>(tmem_base, desc_lo, desc_hi, acc_first);
              umma_commit_2sm(empty_bar + stage);
              if (kb == p.nkb - 1) umma_commit_2sm(tmem_full);
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
          tphase ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue warps 2..5 of both CTAs: own TMEM lanes 32*(warp%4) .. +31 =====
    const int lg = warp & 3;
    double* sb = stg + lg * 32 * 17;
    uint32_t tphase = 0;
    for (int t = cluster_id; t < ntiles; t += nclusters) {
      int I2, J;
      tile2_from_index(p.tile_map, t, I2, J);
      const int64_t row = (int64_t)I2 * 256 + (int64_t)rank * kTile + lg * 32 + lane;
      const double rs_i = (row < p.ns) ? p.rowscale[row] : 0.0;
      for (int pass = 0; pass < npasses; ++pass) {
        const int d_lo = 2 + pass * kLevelsPerPass;
        const int d_hi = min(d_lo + kLevelsPerPass - 1, S + 1);
        mbar_wait(tmem_full, tphase);
        tc_fence_after();
        const bool add = (pass > 0) || (p.accum != 0);
        // whole 128-column block above the diagonal for this CTA's rows? then nothing to write
        const bool any = (int64_t)J * kTile <= (int64_t)I2 * 256 + (int64_t)rank * kTile + kTile - 1;
        if (any) {
          for (int c0 = 0; c0 < kTile; c0 += 16) {
            double sum[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) sum[e] = 0.0;
            for (int d = d_hi; d >= d_lo; --d) {
              int32_t v[16];
              tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(d - d_lo) * kTile + c0, v);
              tmem_ld_wait();
              const double w = scalbn(1.0, -7 * d + 2);
#pragma unroll
              for (int e = 0; e < 16; ++e) sum[e] += w * (double)v[e];
            }
            // A TMEM lane is a row of T, so storing from registers puts the 32 lanes of an instruction 8 ns bytes apart:
            // with T beyond the L2 (ns = 16384: 2.1 GB) the read-modify-write of the later passes / K chunks ran at
            // DRAM-page-miss speed and was 90 % of the kernel (265 ms per launch whatever K).  The 32 x 16 block goes
            // through shared memory instead and is written (and, when accumulating, read) 16 consecutive columns per
            // half warp; only the lower triangle is touched here, gram_mirror_kernel fills the upper one at the end.
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int64_t col = (int64_t)J * kTile + c0 + e;
              sb[lane * 17 + e] = (row < p.ns && col < p.ns) ? sum[e] * (rs_i * __ldg(p.rowscale + col)) : 0.0;
            }
            __syncwarp();
            {
              const int hw = lane >> 4, ce = lane & 15;
              const int64_t row0 = (int64_t)I2 * 256 + (int64_t)rank * kTile + lg * 32;
              const int64_t gcol = (int64_t)J * kTile + c0 + ce;
              double old[16];
#pragma unroll
              for (int rr = 0; rr < 16; ++rr) {
                const int64_t grow = row0 + 2 * rr + hw;
                old[rr] = (add && grow < p.ns && gcol <= grow) ? p.T[grow * p.ns + gcol] : 0.0;
              }
#pragma unroll
              for (int rr = 0; rr < 16; ++rr) {
                const int64_t grow = row0 + 2 * rr + hw;
                if (grow < p.ns && gcol <= grow) p.T[grow * p.ns + gcol] = old[rr] + sb[(2 * rr + hw) * 17 + ce];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tmem_empty);
        tphase ^= 1;
        if constexpr (PUSH) {
          // The tile is final after the last pass: copy this warp's 32 rows (lower triangle only) from the local T
          // (L2-hot; written by the lanes of this warp, ordered by the __syncwarp above) into every peer's staging
          // slot.  Lanes walk a row, so every store instruction is one contiguous 256-byte run over NVLink, and the
          // MMA warp is already working on the next tile (TMEM was released above).
          if (pass == npasses - 1 && any) {
            const int64_t row0 = (int64_t)I2 * 256 + (int64_t)rank * kTile + lg * 32;
            for (int rr = 0; rr < 32; ++rr) {
              const int64_t r = row0 + rr;
              if (r >= p.ns) break;
              const double* srow = p.T + r * p.ns;
#pragma unroll
              for (int e = 0; e < kTile / 32; ++e) {
                const int64_t c = (int64_t)J * kTile + lane + 32 * e;
                if (c <= r) {
                  const double v = __ldcg(srow + c);
                  for (int q = 0; q < pp.nranks; ++q)
                    if (q != pp.rank) pp.slot[q][r * p.ns + c] = v;
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer must not exit (or free TMEM) while the leader still uses its smem / TMEM
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

template <int S>
__global__ void __launch_bounds__(kThreads, 1)
    gram_tc2_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, GramTcParams p) {
  gram_tc2_body<S, false>(tmapA, tmapB, p, GramPushParams{});
}

// last K chunk of the fused Gram + exchange path: same kernel, finished tiles are also stored to the peers
template <int S>
__global__ void __launch_bounds__(kThreads, 1)
    gram_tc2_push_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                         GramTcParams p, const __grid_constant__ GramPushParams pp) {
  gram_tc2_body<S, true>(tmapA, tmapB, p, pp);
}

template <int S>
static int launch_gram_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GramTcParams& p, int grid, size_t smem,
                           cudaStream_t st, const GramPushParams* push) {
  if (push)
    QTX_CUDA(cudaFuncSetAttribute(gram_tc2_push_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    QTX_CUDA(cudaFuncSetAttribute(gram_tc2_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (push) QTX_CUDA(cudaLaunchKernelEx(&cfg, gram_tc2_push_kernel<S>, tmA, tmB, p, *push));
  else QTX_CUDA(cudaLaunchKernelEx(&cfg, gram_tc2_kernel<S>, tmA, tmB, p));
  count_launch();
  return QTX_OK;
}

static int gram_tc2_dispatch(int s, const CUtensorMap& tmA, const CUtensorMap& tmB, const GramTcParams& p, int grid,
                             size_t smem, cudaStream_t st, const GramPushParams* push) {
  switch (s) {
    case 1: return launch_gram_tc2<1>(tmA, tmB, p, grid, smem, st, push);
    case 2: return launch_gram_tc2<2>(tmA, tmB, p, grid, smem, st, push);
    case 3: return launch_gram_tc2<3>(tmA, tmB, p, grid, smem, st, push);
    case 4: return launch_gram_tc2<4>(tmA, tmB, p, grid, smem, st, push);
    case 5: return launch_gram_tc2<5>(tmA, tmB, p, grid, smem, st, push);
    case 6: return launch_gram_tc2<6>(tmA, tmB, p, grid, smem, st, push);
    case 7: return launch_gram_tc2<7>(tmA, tmB, p, grid, smem, st, push);
    default: return launch_gram_tc2<8>(tmA, tmB, p, grid, smem, st, push);
  }
}

int gram_tc2_launch(int s, const CUtensorMap& tmA, const CUtensorMap& tmB, const GramTcParams& p, int grid, size_t smem,
                    cudaStream_t st) {
  return gram_tc2_dispatch(s, tmA, tmB, p, grid, smem, st, nullptr);
}

// last K chunk of qtx_gram_push: slots[q] = staging slot of this rank's partial on rank q (device pointers)
int gram_tc2_launch_push(int s, const CUtensorMap& tmA, const CUtensorMap& tmB, const GramTcParams& p, int grid,
                         size_t smem, cudaStream_t st, int nranks, int rank, double* const* slots) {
  GramPushParams pp{};
  pp.nranks = nranks;
  pp.rank = rank;
  for (int q = 0; q < nranks; ++q) pp.slot[q] = slots[q];
  return gram_tc2_dispatch(s, tmA, tmB, p, grid, smem, st, &pp);
}

}  // namespace qtx
