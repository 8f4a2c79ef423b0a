// Self-test kernels for tests/native/cuda_emu.h (written in CUDA style, run on the CPU): block reduction through
// shuffles and shared memory, ballot-based stream compaction with dynamic shared memory, atomics across blocks.
#include "cuda_emu.h"

__global__ void block_reduce_kernel(const double* __restrict__ x, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// keeps the positive entries of each block's segment, in order (ballot + popc prefix, staging in dynamic smem)
__global__ void compact_kernel(const int* __restrict__ x, int n, int* __restrict__ out, int* __restrict__ count) {
  QTX_DYN_SMEM(int, stage);
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = i < n ? x[i] : 0;
  const unsigned bal = __ballot_sync(FULL, v > 0);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_cnt[w];
  if (v > 0) stage[before + __popc(bal & ((1u << lane) - 1u))] = v;
  __syncthreads();
  int total = 0;
  for (unsigned w = 0; w < (blockDim.x >> 5); ++w) total += warp_cnt[w];
  if (threadIdx.x == 0) base = atomicAdd(count, total);
  __syncthreads();
  for (int k = threadIdx.x; k < total; k += blockDim.x) out[base + k] = stage[k];
}

__global__ void shuffle_kernel(int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  out[threadIdx.x] = __shfl_sync(FULL, lane * 10, 3) + __shfl_down_sync(FULL, lane, 1) + __shfl_xor_sync(FULL, lane, 1);
  __syncwarp();
}

extern "C" {
void run_block_reduce(const double* x, int n, unsigned grid, unsigned block, double* out) {
  *out = 0.0;
  emu_launch(grid, block, [&] { block_reduce_kernel(x, n, out); });
}
void run_compact(const int* x, int n, unsigned block, int* out, int* count) {
  *count = 0;
  emu_launch((unsigned)((n + block - 1) / block), block, [&] { compact_kernel(x, n, out, count); }, block * sizeof(int));
}
void run_shuffle(int* out) { emu_launch(1, 64, [&] { shuffle_kernel(out); }); }
}
