// Minimal host emulation of the CUDA execution model for tests (TEST INFRASTRUCTURE ONLY): one std::thread per CUDA
// thread of a block, blocks run one after the other on the same threads, __syncthreads / __shfl_xor_sync built on std::barrier.  Enough to
// run the small reduction kernels of csrc/pinv_rational.cu unmodified on the CPU (tests/test_pinv_rational_emu_cpu.py).
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static  // blocks run sequentially, so one static instance is the block's shared memory

struct emu_uint3 {
  unsigned x = 0, y = 0, z = 0;
};
struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
inline thread_local emu_uint3 threadIdx, blockIdx;
inline thread_local emu_dim3 blockDim, gridDim;

struct EmuBlock {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> slot;
};
inline thread_local EmuBlock* emu_block = nullptr;

inline void __syncthreads() { emu_block->block_bar->arrive_and_wait(); }

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "emulated shuffle moves at most 8 bytes");
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  emu_block->slot[tid] = bits;
  emu_block->warp_bar[warp]->arrive_and_wait();
  const uint64_t other = emu_block->slot[(warp << 5) | (lane ^ (unsigned)lane_mask)];
  emu_block->warp_bar[warp]->arrive_and_wait();
  T r;
  std::memcpy(&r, &other, sizeof(T));
  return r;
}

// the other warp primitives, on the same exchange slots
template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "emulated shuffle moves at most 8 bytes");
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  emu_block->slot[tid] = bits;
  emu_block->warp_bar[warp]->arrive_and_wait();
  const uint64_t other = emu_block->slot[(warp << 5) | ((unsigned)src_lane & 31u)];
  emu_block->warp_bar[warp]->arrive_and_wait();
  T r;
  std::memcpy(&r, &other, sizeof(T));
  return r;
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
  const unsigned lane = threadIdx.x & 31;
  const T r = __shfl_sync(mask, v, (int)((lane + delta) & 31u));
  return lane + delta < 32 ? r : v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  emu_block->slot[tid] = pred ? 1u : 0u;
  emu_block->warp_bar[warp]->arrive_and_wait();
  unsigned bal = 0;
  const unsigned lanes = (warp + 1) * 32 <= blockDim.x ? 32 : blockDim.x - warp * 32;
  for (unsigned l = 0; l < lanes; ++l) bal |= (unsigned)(emu_block->slot[(warp << 5) | l] & 1u) << l;
  emu_block->warp_bar[warp]->arrive_and_wait();
  return bal;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu_block->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
template <typename T>
inline T atomicAdd(T* addr, T val) {
  return std::atomic_ref<T>(*addr).fetch_add(val, std::memory_order_relaxed);
}

// dynamic shared memory: `extern __shared__ T name[];` cannot be emulated by a macro, so kernels that want to run
// here declare it as  QTX_DYN_SMEM(T, name);  (nvcc: extern __shared__ T name[];  here: the launch's buffer)
inline thread_local unsigned char* emu_dyn_smem = nullptr;
#define QTX_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu_dyn_smem)

// kernel<<<grid, block>>>(args...)  ->  emu_launch(grid, block, [&] { kernel(args...); });
// `block` threads are started once and walk through the blocks of the grid together (a barrier between blocks keeps
// the static "shared memory" of one block from being touched by the next).
template <typename F>
inline void emu_launch(unsigned grid, unsigned block, F body, size_t dyn_smem_bytes = 0) {
  std::vector<unsigned char> dyn(dyn_smem_bytes + 16);
  EmuBlock blk;
  blk.block_bar = std::make_unique<std::barrier<>>(block);
  for (unsigned w = 0; w < (block + 31) / 32; ++w) {
    const unsigned lanes = (w + 1) * 32 <= block ? 32 : block - w * 32;
    blk.warp_bar.push_back(std::make_unique<std::barrier<>>(lanes));
  }
  blk.slot.assign(block, 0);
  std::barrier<> between(block);
  std::vector<std::thread> threads;
  threads.reserve(block);
  for (unsigned t = 0; t < block; ++t)
    threads.emplace_back([&, t] {
      threadIdx.x = t;
      blockDim.x = block;
      gridDim.x = grid;
      emu_block = &blk;
      emu_dyn_smem = dyn.data();
      for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        body();
        between.arrive_and_wait();
      }
    });
  for (auto& th : threads) th.join();
}

// the few CUDA library names the kernels use
constexpr unsigned FULL = 0xffffffffu;
template <typename T>
inline T warp_sum(T v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
struct cuDoubleComplex {
  double x, y;
};
inline cuDoubleComplex make_cuDoubleComplex(double re, double im) { return {re, im}; }
inline double rsqrt(double v) { return 1.0 / std::sqrt(v); }
