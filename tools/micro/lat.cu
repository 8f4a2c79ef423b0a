// Latency microbenchmarks (dev tool): dependent DFMA / DMUL chain, F2F conversions, MUFU.RCP, LDS round trip,
// __syncthreads at 256 threads, STS->BAR->LDS hand-over.  nvcc -arch=sm_100a -O3 lat.cu -o lat && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[256];
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  const int N = 512;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-12);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  // 2. dependent DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  // 3. f64 -> f32 -> rcp -> f64 chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) x = (double)__frcp_rn((float)x) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  // 4. barrier alone
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  // 5. STS -> BAR -> LDS (other thread's value) -> DFMA hand-over
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) {
    sm[threadIdx.x] = x;
    __syncthreads();
    x = fma(sm[(threadIdx.x + 33) & 255], y, 1e-12);
    __syncthreads();
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  // 6. IEEE double division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) x = 1.0 / x + 0.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 64);
  for (int threads : {32, 256}) {
    k<<<1, threads>>>(out, cyc, 1.0); k<<<1, threads>>>(out, cyc, 1.0);
    long long h[6]; cudaMemcpy(h, cyc, 48, cudaMemcpyDeviceToHost);
    printf("threads %d: DFMA dep %.1f cyc, DMUL dep %.1f, f64->f32 rcp ->f64 (+DADD) %.1f, BAR %.1f, STS-BAR-LDS-DFMA-BAR %.1f, 1/x+c %.1f\n",
           threads, h[0] / 512.0, h[1] / 512.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0);
  }
  return 0;
}
