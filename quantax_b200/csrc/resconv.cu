// ResConv (quantax/model/conv_nets.py:26-183): batched forward, per-sample log-derivative
// Jacobian, and the model-agnostic Metropolis propose / accept kernels used when a state has no
// local-update path (quantax/state/variational.py:383-384: full forward per proposal).
//
// Layout: activations [ns, C, Lx*Ly] in model dtype, weights in the reference's flat parameter
// order (conv1.weight [C,Cin,kh,kw], conv1.bias [C], conv2.weight, conv2.bias per block; the last
// conv has no bias).  A convolution layer is an implicit GEMM over (sample, position) x out-channel
// with K = Cin*kh*kw: the CTA stages a channel chunk of whole (circularly padded) images in shared
// memory, so the kh*kw taps re-read shared memory instead of L2, and each thread keeps a
// 4 x 8 (position x out-channel) register tile.  Arithmetic stays in the model dtype (float32 by
// default: the 1e-5 parity bar rules out single-pass TF32), so these kernels are bound by the FP32
// FMA pipe, not by HBM.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace qtx {

// tensor-core forward (resconv_tc.cu)
bool resconv_tc_supported(int C, int lx, int ly, int kh, int kw);
size_t resconv_tc_workspace(int64_t ns, int nblocks, int C, int lx, int ly, int grad);
bool resconv_tc_backward_supported(int C, int lx, int ly, int kh, int kw);
int resconv_tc_backward(int nblocks, int C, int lx, int ly, const float* params, int64_t ns, const float* X,
                        const float* Hs, const float* seed, void* out, int out_f64, int64_t ld, void* ws, size_t ws_bytes,
                        const float** raw_grad, cudaStream_t st);
int resconv_tc_forward(int nblocks, int C, int lx, int ly, const float* params, const int8_t* spins, int64_t ns,
                       float* X, float* Hs, int save_all, void* ws, size_t ws_bytes, const float** x_final,
                       int* x_final_planes, const long long* ns_dev, cudaStream_t st);

constexpr int kConvThreads = 256;
constexpr int kOT = 32;   // out-channel tile per CTA
constexpr int kTN = 8;    // out channels per thread
constexpr int kTM = 4;    // positions per thread
constexpr int kCK = 8;    // in-channel chunk
constexpr int kPT = 256;  // positions per CTA (64 position groups x kTM)

template <typename T>
__device__ __forceinline__ T gelu_f(T x) {
  const T u = T(0.7978845608028654) * (x + T(0.044715) * x * x * x);
  return T(0.5) * x * (T(1) + tanh(u));
}
template <typename T>
__device__ __forceinline__ T gelu_grad_f(T x) {
  const T u = T(0.7978845608028654) * (x + T(0.044715) * x * x * x);
  const T t = tanh(u);
  const T du = T(0.7978845608028654) * (T(1) + T(3 * 0.044715) * x * x);
  return T(0.5) * (T(1) + t) + T(0.5) * x * (T(1) - t * t) * du;
}

template <typename T>
struct ConvParams {
  const T* in;      // [ns, cin, N]
  const T* w;       // [cout, cin, kh, kw]  (or pre-transposed for the backward-data pass)
  const T* w_ctco;  // the same weights repacked as [cin, kh*kw, cout] (fast kernels, coalesced staging)
  const T* bias;    // [cout] or null
  const T* res;     // residual added after the multiply, [ns, res_ch, N], res_ch in {0, 1, cout}
  const T* mul;     // if non-null: out = (conv + bias) * mul_scale * gelu'(mul_alpha * mul[s,o,r]) + res
  T* out;           // [ns, cout, N]
  int64_t ns;
  int cin, cout, lx, ly, kh, kw, res_ch;
  int in_act;       // 0: f(x) = alpha*x ; 1: f(x) = gelu(alpha*x)
  T alpha, mul_alpha, mul_scale;
  int s_tile;       // samples per CTA
};

template <typename T>
__global__ void __launch_bounds__(kConvThreads) conv_circ_kernel(ConvParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.lx * p.ly, ph = (p.kh - 1) / 2, pw = (p.kw - 1) / 2;
  const int Hp = p.lx + p.kh - 1, Wp = p.ly + p.kw - 1, HW = Hp * Wp;
  const int taps = p.kh * p.kw;
  T* act_s = reinterpret_cast<T*>(smem_raw);                 // [kCK][s_tile][HW]
  T* w_s = act_s + (size_t)kCK * p.s_tile * HW;              // [kCK][taps][kOT]
  const int tid = threadIdx.x;
  const int og = tid >> 6, pg = tid & 63;                    // 4 channel groups x 64 position groups
  const int64_t s0 = (int64_t)blockIdx.x * p.s_tile;
  const int o0 = blockIdx.y * kOT;
  const int pos0 = blockIdx.z * kPT;                         // position tile inside the sample group
  const int P = p.s_tile * N;

  int base[kTM];
  bool pvalid[kTM];
#pragma unroll
  for (int t = 0; t < kTM; ++t) {
    int pos = pos0 + pg + 64 * t;
    pvalid[t] = pos < P && (s0 + pos / N) < p.ns;
    int sl = pvalid[t] ? pos / N : 0, r = pvalid[t] ? pos % N : 0;
    base[t] = sl * HW + (r / p.ly) * Wp + (r % p.ly);
  }
  T acc[kTM][kTN];
#pragma unroll
  for (int t = 0; t < kTM; ++t)
#pragma unroll
    for (int n = 0; n < kTN; ++n) acc[t][n] = 0;

  for (int c0 = 0; c0 < p.cin; c0 += kCK) {
    __syncthreads();
    // stage activations (with the input transform and the circular halo)
    const int nact = kCK * p.s_tile * HW;
    for (int e = tid; e < nact; e += kConvThreads) {
      int ck = e / (p.s_tile * HW), rem = e % (p.s_tile * HW);
      int sl = rem / HW, q = rem % HW;
      int hp = q / Wp, wp = q % Wp;
      int h = hp - ph, w = wp - pw;
      h += (h < 0) ? p.lx : 0; h -= (h >= p.lx) ? p.lx : 0;
      w += (w < 0) ? p.ly : 0; w -= (w >= p.ly) ? p.ly : 0;
      T v = 0;
      int c = c0 + ck;
      if (c < p.cin && s0 + sl < p.ns) {
        v = p.alpha * p.in[((s0 + sl) * p.cin + c) * N + h * p.ly + w];
        if (p.in_act == 1) v = gelu_f(v);
      }
      act_s[e] = v;
    }
    // stage weights as [ck][tap][o]
    const int nw = kCK * taps * kOT;
    for (int e = tid; e < nw; e += kConvThreads) {
      int o = e % kOT, tap = (e / kOT) % taps, ck = e / (kOT * taps);
      int c = c0 + ck;
      T v = 0;
      if (c < p.cin && o0 + o < p.cout) v = p.w[((size_t)(o0 + o) * p.cin + c) * taps + tap];
      w_s[e] = v;
    }
    __syncthreads();
    for (int ck = 0; ck < kCK; ++ck) {
      const T* a_c = act_s + (size_t)ck * p.s_tile * HW;
      const T* w_c = w_s + (size_t)ck * taps * kOT + og * kTN;
      for (int dy = 0; dy < p.kh; ++dy)
        for (int dx = 0; dx < p.kw; ++dx) {
          const int off = dy * Wp + dx;
          T a[kTM], wv[kTN];
#pragma unroll
          for (int t = 0; t < kTM; ++t) a[t] = a_c[base[t] + off];
          const T* wr = w_c + (dy * p.kw + dx) * kOT;
#pragma unroll
          for (int n = 0; n < kTN; ++n) wv[n] = wr[n];
#pragma unroll
          for (int t = 0; t < kTM; ++t)
#pragma unroll
            for (int n = 0; n < kTN; ++n) acc[t][n] += a[t] * wv[n];
        }
    }
  }
#pragma unroll
  for (int t = 0; t < kTM; ++t) {
    if (!pvalid[t]) continue;
    int pos = pos0 + pg + 64 * t;
    int64_t s = s0 + pos / N;
    int r = pos % N;
#pragma unroll
    for (int n = 0; n < kTN; ++n) {
      int o = o0 + og * kTN + n;
      if (o >= p.cout) continue;
      T v = acc[t][n];
      if (p.bias) v += p.bias[o];
      const int64_t oi = (s * p.cout + o) * N + r;
      if (p.mul) v = v * (p.mul_scale * gelu_grad_f(p.mul_alpha * p.mul[oi]));
      if (p.res_ch == p.cout) v += p.res[oi];
      else if (p.res_ch == 1) v += p.res[s * N + r];
      p.out[oi] = v;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// fast path: compile-time kernel size and out-channel register tile.  Per (channel, tap) step a
// thread issues 4 scalar LDS (activations), TN/4 LDS.128 (weights) and 4*TN FFMA, so ~85% of the
// inner-loop instructions are FMAs; staging indices are computed once per thread.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 4 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T, int KH, int KW, int TN>
__global__ void __launch_bounds__(kConvThreads, 2) conv_circ_fast_kernel(ConvParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int OT = 4 * TN, TAPS = KH * KW, PH = (KH - 1) / 2, PW = (KW - 1) / 2;
  constexpr int WSZ = kCK * TAPS * OT;
  const int N = p.lx * p.ly;
  const int Hp = p.lx + KH - 1, Wp = p.ly + KW - 1, HW = Hp * Wp;
  const int HWs = (HW + 3) & ~3;                               // image stride in smem
  const int nimg = kCK * p.s_tile;
  const int RSZ = (nimg * N + 3) & ~3;                         // raw (un-padded) chunk size
  T* act_s = reinterpret_cast<T*>(smem_raw);                   // [kCK][s_tile][HWs]   transformed + halo
  T* w_s = act_s + (size_t)nimg * HWs;                         // [2][kCK][TAPS][OT]   double buffered
  T* raw_s = w_s + 2 * WSZ;                                    // [2][kCK][s_tile][N]  double buffered
  int* qsrc_s = reinterpret_cast<int*>(raw_s + 2 * (size_t)RSZ);
  const int tid = threadIdx.x;
  const int og = tid >> 6, pg = tid & 63;
  const int64_t s0 = (int64_t)blockIdx.x * p.s_tile;
  const int o0 = blockIdx.y * OT;
  const int pos0 = blockIdx.z * kPT;
  const int P = p.s_tile * N;

  int base[kTM];
  bool pvalid[kTM];
#pragma unroll
  for (int t = 0; t < kTM; ++t) {
    int pos = pos0 + pg + 64 * t;
    pvalid[t] = pos < P && (s0 + pos / N) < p.ns;
    int sl = pvalid[t] ? pos / N : 0, r = pvalid[t] ? pos % N : 0;
    base[t] = sl * HWs + (r / p.ly) * Wp + (r % p.ly);
  }
  // staging map: padded position q of an image -> source position, -1 outside
  for (int q = tid; q < HWs; q += kConvThreads) {
    int hp = q / Wp, wp = q - hp * Wp;
    int h = hp - PH, w = wp - PW;
    h += (h < 0) ? p.lx : 0; h -= (h >= p.lx) ? p.lx : 0;
    w += (w < 0) ? p.ly : 0; w -= (w >= p.ly) ? p.ly : 0;
    qsrc_s[q] = (q < HW) ? h * p.ly + w : -1;
  }
  T acc[kTM][TN];
#pragma unroll
  for (int t = 0; t < kTM; ++t)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[t][n] = 0;

  // asynchronous (cp.async) fetch of the raw activations and the repacked weights of one channel chunk;
  // issued one chunk ahead so that global-memory latency overlaps with the FMA loop of the current chunk.
  auto fetch = [&](int c0, int buf) {
    T* rdst = raw_s + (size_t)buf * RSZ;
    for (int e = tid; e < nimg * N; e += kConvThreads) {
      const int img = e / N, r = e - img * N;
      const int ck = img / p.s_tile, sl = img - ck * p.s_tile;
      const bool ok = c0 + ck < p.cin && s0 + sl < p.ns;
      const T* src = ok ? p.in + ((s0 + sl) * p.cin + c0 + ck) * N + r : p.in;
      if constexpr (sizeof(T) == 4) cp_async4(rdst + e, src, ok);
      else cp_async8(rdst + e, src, ok);
    }
    T* wdst = w_s + (size_t)buf * WSZ;
    for (int e = tid; e < WSZ; e += kConvThreads) {
      const int o = e % OT, row = e / OT;  // row = ck * TAPS + tap
      const bool ok = c0 + row / TAPS < p.cin && o0 + o < p.cout;
      const T* src = ok ? p.w_ctco + ((size_t)c0 * TAPS + row) * p.cout + o0 + o : p.w_ctco;
      if constexpr (sizeof(T) == 4) cp_async4(wdst + e, src, ok);
      else cp_async8(wdst + e, src, ok);
    }
    cp_async_commit();
  };
  fetch(0, 0);
  int buf = 0;
  for (int c0 = 0; c0 < p.cin; c0 += kCK, buf ^= 1) {
    cp_async_wait_all();
    __syncthreads();  // chunk `buf` has landed; the previous FMA loop is done with act_s
    const T* rsrc = raw_s + (size_t)buf * RSZ;
    for (int e = tid; e < nimg * HWs; e += kConvThreads) {
      const int img = e / HWs, q = e - img * HWs;
      const int src = qsrc_s[q];
      if (src >= 0) {
        T v = p.alpha * rsrc[img * N + src];
        if (p.in_act == 1) v = gelu_f(v);
        act_s[e] = v;
      }
    }
    __syncthreads();
    if (c0 + kCK < p.cin) fetch(c0 + kCK, buf ^ 1);
    const T* w_b = w_s + (size_t)buf * WSZ;
#pragma unroll 2
    for (int ck = 0; ck < kCK; ++ck) {
      const T* a_c = act_s + (size_t)ck * p.s_tile * HWs;
      const T* w_c = w_b + (size_t)ck * TAPS * OT + og * TN;
#pragma unroll
      for (int dy = 0; dy < KH; ++dy)
#pragma unroll
        for (int dx = 0; dx < KW; ++dx) {
          const int off = dy * Wp + dx;
          T a[kTM], wv[TN];
#pragma unroll
          for (int t = 0; t < kTM; ++t) a[t] = a_c[base[t] + off];
          const T* wr = w_c + (dy * KW + dx) * OT;
          if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int n = 0; n < TN; n += 4) {
              float4 q4 = *reinterpret_cast<const float4*>(wr + n);
              wv[n] = q4.x; wv[n + 1] = q4.y; wv[n + 2] = q4.z; wv[n + 3] = q4.w;
            }
          } else {
#pragma unroll
            for (int n = 0; n < TN; n += 2) {
              double2 q2 = *reinterpret_cast<const double2*>(wr + n);
              wv[n] = q2.x; wv[n + 1] = q2.y;
            }
          }
#pragma unroll
          for (int t = 0; t < kTM; ++t)
#pragma unroll
            for (int n = 0; n < TN; ++n) acc[t][n] += a[t] * wv[n];
        }
    }
  }
#pragma unroll
  for (int t = 0; t < kTM; ++t) {
    if (!pvalid[t]) continue;
    int pos = pos0 + pg + 64 * t;
    int64_t s = s0 + pos / N;
    int r = pos % N;
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      int o = o0 + og * TN + n;
      if (o >= p.cout) continue;
      T v = acc[t][n];
      if (p.bias) v += p.bias[o];
      const int64_t oi = (s * p.cout + o) * N + r;
      if (p.mul) v = v * (p.mul_scale * gelu_grad_f(p.mul_alpha * p.mul[oi]));
      if (p.res_ch == p.cout) v += p.res[oi];
      else if (p.res_ch == 1) v += p.res[s * N + r];
      p.out[oi] = v;
    }
  }
}

template <typename T, int KH, int KW, int TN>
static int launch_conv_fast(ConvParams<T> p, cudaStream_t st) {
  const int N = p.lx * p.ly;
  const int Hp = p.lx + KH - 1, Wp = p.ly + KW - 1, HWs = (Hp * Wp + 3) & ~3;
  constexpr int OT = 4 * TN;
  const int nimg = kCK * p.s_tile, RSZ = (nimg * N + 3) & ~3;
  size_t smem = ((size_t)nimg * HWs + 2 * (size_t)kCK * KH * KW * OT + 2 * (size_t)RSZ) * sizeof(T) + (size_t)HWs * sizeof(int);
  auto k = conv_circ_fast_kernel<T, KH, KW, TN>;
  if (smem > 48 * 1024) QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int P = p.s_tile * N;
  dim3 grid((unsigned)((p.ns + p.s_tile - 1) / p.s_tile), (unsigned)((p.cout + OT - 1) / OT),
            (unsigned)((P + kPT - 1) / kPT));
  k<<<grid, kConvThreads, smem, st>>>(p);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

template <typename T, int KH, int KW>
static int launch_conv_fast_tn(ConvParams<T> p, cudaStream_t st) {
  // out-channel tile 4*TN: pick the TN in {8, 12, 16} with the least padded work, larger tile on ties
  int best = 8;
  double best_eff = 0;
  for (int tn : {8, 12, 16}) {
    int ot = 4 * tn, tiles = (p.cout + ot - 1) / ot;
    double eff = (double)p.cout / (tiles * ot) * (tn == 8 ? 0.93 : (tn == 12 ? 0.97 : 1.0));
    if (eff > best_eff + 1e-9) { best_eff = eff; best = tn; }
  }
  if (best == 8) return launch_conv_fast<T, KH, KW, 8>(p, st);
  if (best == 12) return launch_conv_fast<T, KH, KW, 12>(p, st);
  return launch_conv_fast<T, KH, KW, 16>(p, st);
}

template <typename T>
static int launch_conv(ConvParams<T> p, cudaStream_t st) {
  const int N = p.lx * p.ly;
  p.s_tile = kPT / N > 0 ? kPT / N : 1;
  const int Hp = p.lx + p.kh - 1, Wp = p.ly + p.kw - 1;
  if (Hp * Wp <= 384 && !getenv("QTX_CONV_GENERIC")) {
    if (p.kh == 3 && p.kw == 3) return launch_conv_fast_tn<T, 3, 3>(p, st);
    if (p.kh == 1 && p.kw == 3) return launch_conv_fast_tn<T, 1, 3>(p, st);
  }
  size_t smem = ((size_t)kCK * p.s_tile * Hp * Wp + (size_t)kCK * p.kh * p.kw * kOT) * sizeof(T);
  QTX_REQUIRE(smem <= 200 * 1024, QTX_ERR_UNSUPPORTED, "resconv: lattice too large for the conv tile (%zu B smem)", smem);
  auto k = conv_circ_kernel<T>;
  if (smem > 48 * 1024) QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int P = p.s_tile * N;
  dim3 grid((unsigned)((p.ns + p.s_tile - 1) / p.s_tile), (unsigned)((p.cout + kOT - 1) / kOT),
            (unsigned)((P + kPT - 1) / kPT));
  k<<<grid, kConvThreads, smem, st>>>(p);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void spins_to_act_kernel(const int8_t* __restrict__ s, int64_t n, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (T)s[i];
}

// out[c][tap][o] = w[o][c][tap]            (flip == 0: forward conv, cin = c, cout = o)
// out[o][tap'][c] = w[o][c][flip(tap')]     (flip == 1: backward-data conv, cin' = o, cout' = c)
template <typename T>
__global__ void weight_repack_kernel(const T* __restrict__ w, int cout, int cin, int kh, int kw, int flip,
                                     T* __restrict__ out) {
  const int taps = kh * kw, n = cout * cin * taps;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int tap = e % taps, c = (e / taps) % cin, o = e / (taps * cin);
    if (!flip) {
      out[((size_t)c * taps + tap) * cout + o] = w[e];
    } else {
      int dy = tap / kw, dx = tap % kw;
      int tf = (kh - 1 - dy) * kw + (kw - 1 - dx);
      out[((size_t)o * taps + tf) * cin + c] = w[e];
    }
  }
}

// wT[c][o][kh-1-dy][kw-1-dx] = w[o][c][dy][dx]   (backward-data pass = conv with these weights)
template <typename T>
__global__ void weight_transpose_flip_kernel(const T* __restrict__ w, int cout, int cin, int kh, int kw,
                                             T* __restrict__ wT) {
  const int taps = kh * kw, n = cout * cin * taps;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int tap = e % taps, c = (e / taps) % cin, o = e / (taps * cin);
    int dy = tap / kw, dx = tap % kw;
    wT[((size_t)c * cout + o) * taps + (kh - 1 - dy) * kw + (kw - 1 - dx)] = w[e];
  }
}

// final layer (conv_nets.py:167-173, nn/activation.py:7-14,26-32, nn/conv.py:61-68): one CTA per sample.
//   z = x / sqrt(nblocks+1); m = max|z|; sig = exp(z-m) | sinh-plus-one form; a_r = mean_c sig;
//   significand = sum_r a_r * c1, exponent = m + log(1/N)           (all in model dtype)
// If dz != null also writes the backward seed dz = dsig / sum(sig) / sqrt(nblocks+1).
template <typename T>
__global__ void __launch_bounds__(256) resconv_final_kernel(const T* __restrict__ x, int64_t ns, int C, int N,
                                                            T inv_norm, int final_act, double* __restrict__ sig_out,
                                                            double* __restrict__ exp_out, T* __restrict__ dz, int planes,
                                                            const long long* __restrict__ ns_dev) {
  __shared__ T red[32];
  __shared__ T bc;
  const int64_t s = blockIdx.x;
  if (ns_dev && s >= *ns_dev) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int CN = C * N;
  // planes > 0: x is the planar residual stream [ns, planes, N, 8] of the tensor-core forward (resconv_tc.cu)
  const T* xs = x + s * (planes > 0 ? (int64_t)planes * N * 8 : (int64_t)CN);
  auto at = [&](int c, int r) -> T { return planes > 0 ? xs[((c >> 3) * N + r) * 8 + (c & 7)] : xs[c * N + r]; };
  T m = 0;
  if (planes > 0) {
    if constexpr (sizeof(T) == 4) {
      // 16-byte loads: 4 consecutive channels of one pixel
      const float4* x4 = reinterpret_cast<const float4*>(xs);
      for (int e4 = tid; e4 < planes * N * 2; e4 += blockDim.x) {
        const int c0 = (e4 / (N * 2)) * 8 + (e4 & 1) * 4;
        const float4 q = x4[e4];
        if (c0 + 0 < C) m = fmax(m, fabs(q.x * inv_norm));
        if (c0 + 1 < C) m = fmax(m, fabs(q.y * inv_norm));
        if (c0 + 2 < C) m = fmax(m, fabs(q.z * inv_norm));
        if (c0 + 3 < C) m = fmax(m, fabs(q.w * inv_norm));
      }
    } else {
      for (int e = tid; e < planes * N * 8; e += blockDim.x)
        if (((e / (N * 8)) * 8 + (e & 7)) < C) m = fmax(m, fabs(xs[e] * inv_norm));
    }
  } else {
    for (int e = tid; e < CN; e += blockDim.x) m = fmax(m, fabs(xs[e] * inv_norm));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    T mm = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
    bc = mm;
  }
  __syncthreads();
  m = bc;
  const T em = exp(-m);
  // thread owns positions r = tid, tid+blockDim, ...: channel mean first, then the sum over positions
  T part = 0, tot = 0;
  if (planes > 0) {
    // planar stream: sum_r mean_c sig = (sum over all valid elements) / C, read contiguously
    auto term = [&](T xv) -> T {
      T z = xv * inv_norm;
      T sg = exp(z - m);
      if (final_act == 1) sg = (sg - exp(-z - m)) / T(2) + em;
      return sg;
    };
    if constexpr (sizeof(T) == 4) {
      const float4* x4 = reinterpret_cast<const float4*>(xs);
      T t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      for (int e4 = tid; e4 < planes * N * 2; e4 += blockDim.x) {
        const int c0 = (e4 / (N * 2)) * 8 + (e4 & 1) * 4;
        const float4 q = x4[e4];
        if (c0 + 0 < C) t0 += term(q.x);
        if (c0 + 1 < C) t1 += term(q.y);
        if (c0 + 2 < C) t2 += term(q.z);
        if (c0 + 3 < C) t3 += term(q.w);
      }
      tot = (t0 + t1) + (t2 + t3);
    } else {
      for (int e = tid; e < planes * N * 8; e += blockDim.x) {
        if (((e / (N * 8)) * 8 + (e & 7)) >= C) continue;
        tot += term(xs[e]);
      }
    }
    part = tot / (T)C;
  } else {
    for (int r = tid; r < N; r += blockDim.x) {
      T a = 0;
      for (int c = 0; c < C; ++c) {
        T z = at(c, r) * inv_norm;
        T sg = exp(z - m);
        if (final_act == 1) sg = (sg - exp(-z - m)) / T(2) + em;
        a += sg;
      }
      tot += a;
      part += a / (T)C;
    }
  }
  __syncthreads();
  T p1 = warp_sum(part), p2 = warp_sum(tot);
  if (lane == 0) { red[warp] = p1; red[warp + 16] = p2; }
  __syncthreads();
  if (tid == 0) {
    T a = 0, b = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += red[w]; b += red[w + 16]; }
    const T ch = T(1) / (T)N;          // character of sector 0 (symmetry.py:391)
    const T ech = log(ch);             // ScaleArray.from_value(character).normalize() (big_array.py:442-451)
    const T c1 = ch * exp(T(0) - ech);
    if (sig_out) sig_out[s] = (double)(a * c1);
    if (exp_out) exp_out[s] = (double)(m + ech);
    bc = b;
  }
  if (dz) {
    __syncthreads();
    const T total = bc;
    T* d = dz + s * C * N;
    for (int e = tid; e < CN; e += blockDim.x) {
      T z = xs[e] * inv_norm;
      T ds = exp(z - m);
      if (final_act == 1) ds = (ds + exp(-z - m)) / T(2);
      d[e] = ds / total * inv_norm;
    }
  }
}

// The same final layer for the forward-only tensor-core tower (planar float32 stream, no seed output) in ONE pass over
// the sample: every thread keeps a running max m of |z| and the sums of exp(z - m), exp(-z - m) relative to it
// (rescaled when the max moves), the block merges the (m, A, B) triples.  The two-pass kernel above read the 98 KB of
// a config E sample twice; this kernel runs once per forward of the sweep and of Oloc.
__global__ void __launch_bounds__(256) resconv_final_planar_kernel(const float* __restrict__ x, int64_t ns, int C, int N,
                                                                   float inv_norm, int final_act,
                                                                   double* __restrict__ sig_out, double* __restrict__ exp_out,
                                                                   int planes, const long long* __restrict__ ns_dev) {
  __shared__ float red[3][8];
  const int64_t s = blockIdx.x;
  if (ns_dev && s >= *ns_dev) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* x4 = reinterpret_cast<const float4*>(x + s * (int64_t)planes * N * 8);
  float m = 0.f, A = 0.f, B = 0.f;
  for (int e4 = tid; e4 < planes * N * 2; e4 += blockDim.x) {
    const int c0 = (e4 / (N * 2)) * 8 + (e4 & 1) * 4;
    if (c0 >= C) continue;
    const float4 q = x4[e4];
    float z[4] = {q.x * inv_norm, q.y * inv_norm, q.z * inv_norm, q.w * inv_norm};
    const int nv = min(4, C - c0);
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nv) mx = fmaxf(mx, fabsf(z[i]));
    if (mx > m) {
      const float r = expf(m - mx);
      A *= r; B *= r; m = mx;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nv) {
        A += expf(z[i] - m);
        if (final_act == 1) B += expf(-z[i] - m);
      }
  }
  auto merge = [](float& m1, float& a1, float& b1, float m2, float a2, float b2) {
    const float mm = fmaxf(m1, m2);
    const float r1 = expf(m1 - mm), r2 = expf(m2 - mm);
    a1 = a1 * r1 + a2 * r2;
    b1 = b1 * r1 + b2 * r2;
    m1 = mm;
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(FULL, m, o), a2 = __shfl_xor_sync(FULL, A, o), b2 = __shfl_xor_sync(FULL, B, o);
    merge(m, A, B, m2, a2, b2);
  }
  if (lane == 0) { red[0][warp] = m; red[1][warp] = A; red[2][warp] = B; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) merge(m, A, B, red[0][w], red[1][w], red[2][w]);
    float tot = A;
    if (final_act == 1) tot = (A - B) * 0.5f + (float)(C * N) * expf(-m);
    const float a = tot / (float)C;
    const float ch = 1.0f / (float)N;  // character of sector 0 (symmetry.py:391)
    const float ech = logf(ch);        // ScaleArray.from_value(character).normalize() (big_array.py:442-451)
    const float c1 = ch * expf(0.f - ech);
    if (sig_out) sig_out[s] = (double)(a * c1);
    if (exp_out) exp_out[s] = (double)(m + ech);
  }
}

// Complex-output final layer (conv_nets.py:165-173 with out_dtype complex: pair_cpl, nn/activation.py:75-81):
//   z_c = (x_c + i x_{c + C/2}) / sqrt(nblocks+1), cast to complex128; m = max|z|; sig = exp(z - m) | sinh-plus-one
//   a_r = mean_c sig; psi = ScaleArray(sum_r a_r * c1, m + log(1/N)), all in complex128 / float64.
// Backward seeds (variational.py:461-478, real parameters / complex output): with w = dsig / sum(sig),
//   d Re(log psi) -> dz_re[c] = Re w, dz_re[c + C/2] = -Im w;   d Im(log psi) -> dz_im[c] = Im w, dz_im[c + C/2] = Re w
// (both times 1/sqrt(nblocks+1)).  One CTA per sample.
template <typename T>
__global__ void __launch_bounds__(256) resconv_final_cplx_kernel(const T* __restrict__ x, int64_t ns, int C, int N,
                                                                 T inv_norm, int final_act, double2* __restrict__ sig_out,
                                                                 double* __restrict__ exp_out, T* __restrict__ dz_re,
                                                                 T* __restrict__ dz_im, int planes,
                                                                 const long long* __restrict__ ns_dev) {
  __shared__ double red[3][8];
  __shared__ double bc[3];
  const int64_t s = blockIdx.x;
  if (ns_dev && s >= *ns_dev) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C2 = C / 2, CN2 = C2 * N;
  const T* xs = x + s * (planes > 0 ? (int64_t)planes * N * 8 : (int64_t)C * N);
  auto at = [&](int c, int r) -> T { return planes > 0 ? xs[((c >> 3) * N + r) * 8 + (c & 7)] : xs[c * N + r]; };
  double m = 0;
  for (int e = tid; e < CN2; e += blockDim.x) {
    const int c = e / N, r = e - c * N;
    m = fmax(m, hypot((double)(at(c, r) * inv_norm), (double)(at(c + C2, r) * inv_norm)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
  if (lane == 0) red[0][warp] = m;
  __syncthreads();
  if (tid == 0) {
    double mm = red[0][0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) mm = fmax(mm, red[0][w]);
    bc[0] = mm;
  }
  __syncthreads();
  m = bc[0];
  const double em = exp(-m);
  double sre = 0, sim = 0;
  for (int e = tid; e < CN2; e += blockDim.x) {
    const int c = e / N, r = e - c * N;
    const double zr = (double)(at(c, r) * inv_norm), zi = (double)(at(c + C2, r) * inv_norm);
    double sn, cs;
    sincos(zi, &sn, &cs);
    const double ep = exp(zr - m);
    double gr = ep * cs, gi = ep * sn;
    if (final_act == 1) {
      const double en = exp(-zr - m);
      gr = (gr - en * cs) * 0.5 + em;
      gi = (gi + en * sn) * 0.5;
    }
    sre += gr;
    sim += gi;
  }
  __syncthreads();
  sre = warp_sum(sre);
  sim = warp_sum(sim);
  if (lane == 0) { red[1][warp] = sre; red[2][warp] = sim; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += red[1][w]; b += red[2][w]; }
    const double ch = 1.0 / (double)N, ech = log(ch), c1 = ch * exp(0.0 - ech);
    if (sig_out) sig_out[s] = make_double2(a / C2 * c1, b / C2 * c1);
    if (exp_out) exp_out[s] = m + ech;
    bc[1] = a;
    bc[2] = b;
  }
  if (dz_re) {
    __syncthreads();
    const double tr = bc[1], ti = bc[2], den = tr * tr + ti * ti;
    T* dr = dz_re + s * C * N;
    T* di = dz_im + s * C * N;
    for (int e = tid; e < CN2; e += blockDim.x) {
      const int c = e / N, r = e - c * N;
      const double zr = (double)(at(c, r) * inv_norm), zi = (double)(at(c + C2, r) * inv_norm);
      double sn, cs;
      sincos(zi, &sn, &cs);
      const double ep = exp(zr - m);
      double gr = ep * cs, gi = ep * sn;
      if (final_act == 1) {
        const double en = exp(-zr - m);
        gr = (gr + en * cs) * 0.5;
        gi = (gi - en * sn) * 0.5;
      }
      // w = dsig / total
      const double wr = (gr * tr + gi * ti) / den, wi = (gi * tr - gr * ti) / den;
      dr[c * N + r] = (T)(wr * (double)inv_norm);
      dr[(c + C2) * N + r] = (T)(-wi * (double)inv_norm);
      di[c * N + r] = (T)(wi * (double)inv_norm);
      di[(c + C2) * N + r] = (T)(wr * (double)inv_norm);
    }
  }
}

// per-sample weight gradient written straight into the Jacobian row:
//   O[s, col0 + (o*cin + c)*taps + tap] = sum_r delta[s,o,r] * f(alpha * a[s,c,r+tap])
// CTA per (sample, out-channel tile of 32); thread tile 4 (o) x 3 (q = flattened (c, tap)).
template <typename T, typename OutT>
__global__ void __launch_bounds__(192) conv_wgrad_kernel(const T* __restrict__ delta, const T* __restrict__ a, int cin,
                                                         int cout, int lx, int ly, int kh, int kw, T alpha, int in_act,
                                                         OutT* __restrict__ out, int64_t ld, int64_t col0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = lx * ly, ph = (kh - 1) / 2, pw = (kw - 1) / 2;
  const int Hp = lx + kh - 1, Wp = ly + kw - 1, HW = Hp * Wp, taps = kh * kw;
  T* d_s = reinterpret_cast<T*>(smem_raw);   // [32][N]
  T* a_s = d_s + 32 * N;                     // [kCK][HW]
  const int64_t s = blockIdx.x;
  const int o0 = blockIdx.y * 32;
  const int tid = threadIdx.x, og = tid / 24, qg = tid % 24;
  for (int e = tid; e < 32 * N; e += blockDim.x) {
    int o = e / N, r = e % N;
    d_s[e] = (o0 + o < cout) ? delta[(s * cout + o0 + o) * N + r] : T(0);
  }
  const int QC = kCK * taps;  // q's per chunk
  for (int c0 = 0; c0 < cin; c0 += kCK) {
    __syncthreads();
    for (int e = tid; e < kCK * HW; e += blockDim.x) {
      int ck = e / HW, q = e % HW, hp = q / Wp, wp = q % Wp;
      int h = hp - ph, w = wp - pw;
      h += (h < 0) ? lx : 0; h -= (h >= lx) ? lx : 0;
      w += (w < 0) ? ly : 0; w -= (w >= ly) ? ly : 0;
      T v = 0;
      if (c0 + ck < cin) {
        v = alpha * a[(s * cin + c0 + ck) * N + h * ly + w];
        if (in_act == 1) v = gelu_f(v);
      }
      a_s[e] = v;
    }
    __syncthreads();
    for (int qb = qg * 3; qb < QC; qb += 72) {
      int offs[3];
      bool qv[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        int q = qb + j;
        qv[j] = q < QC && (c0 + q / taps) < cin;
        int ck = qv[j] ? q / taps : 0, tap = qv[j] ? q % taps : 0;
        offs[j] = ck * HW + (tap / kw) * Wp + (tap % kw);
      }
      T acc[4][3];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j] = 0;
      for (int h = 0; h < lx; ++h)
        for (int w = 0; w < ly; ++w) {
          const int r = h * ly + w, pos = h * Wp + w;
          T dv[4], av[3];
#pragma unroll
          for (int i = 0; i < 4; ++i) dv[i] = d_s[(og * 4 + i) * N + r];
#pragma unroll
          for (int j = 0; j < 3; ++j) av[j] = a_s[offs[j] + pos];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[i][j] += dv[i] * av[j];
        }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int o = o0 + og * 4 + i;
        if (o >= cout) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (!qv[j]) continue;
          int q = qb + j;
          int c = c0 + q / taps, tap = q % taps;
          out[s * ld + col0 + ((int64_t)o * cin + c) * taps + tap] = (OutT)acc[i][j];
        }
      }
    }
  }
}

// bias gradient: O[s, col0 + o] = sum_r delta[s,o,r]; warp per (sample, o)
template <typename T, typename OutT>
__global__ void conv_bgrad_kernel(const T* __restrict__ delta, int64_t ns, int cout, int N, OutT* __restrict__ out,
                                  int64_t ld, int64_t col0) {
  const int lane = threadIdx.x & 31;
  int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= ns * cout) return;
  int64_t s = wid / cout;
  int o = (int)(wid % cout);
  const T* d = delta + (s * cout + o) * N;
  T acc = 0;
  for (int r = lane; r < N; r += 32) acc += d[r];
  acc = warp_sum(acc);
  if (lane == 0) out[s * ld + col0 + o] = (OutT)acc;
}

// ---------------------------------------------------------------------------------------------
// generic Metropolis propose / accept (models without local updates)
// ---------------------------------------------------------------------------------------------
struct ProposeParams {
  const int8_t* spins;  // [ns, N] current chains
  int8_t* new_spins;    // [ns, N]
  uint8_t* moved;       // [ns]
  int64_t ns;
  int N, kind, max_nb, hop;
  const int32_t* nbr;
  const int32_t* inj_pos;   // [ns] for this step (nullable -> Philox)
  const int32_t* inj_slot;
  uint32_t seed_lo, seed_hi;
  uint64_t step, chain0;
};

__global__ void __launch_bounds__(256) propose_kernel(ProposeParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= p.ns) return;
  const int8_t* sp = p.spins + chain * p.N;
  int8_t* np_ = p.new_spins + chain * p.N;
  int pos, slot = 0;
  if (p.inj_pos) {
    pos = p.inj_pos[chain];
    if (p.kind == QTX_SPIN_EXCHANGE) slot = p.inj_slot[chain];
  } else {
    uint32_t r0 = (uint32_t)(p.chain0 + chain), r1 = (uint32_t)p.step, r2 = (uint32_t)(p.step >> 32), r3 = 0;
    philox4x32_10(r0, r1, r2, r3, p.seed_lo, p.seed_hi);
    if (p.kind == QTX_LOCAL_FLIP) {
      pos = (int)__umulhi(r0, (uint32_t)p.N);
    } else {
      // k-th site holding the hopping particle (same stream mapping as rbm_sweep_kernel)
      int nhop = 0;
      for (int j0 = 0; j0 < p.N; j0 += 32) {
        int j = j0 + lane;
        nhop += __popc(__ballot_sync(FULL, j < p.N && sp[j] == p.hop));
      }
      int k = (int)__umulhi(r0, (uint32_t)nhop);
      pos = 0;
      int seen = 0;
      for (int j0 = 0; j0 < p.N; j0 += 32) {
        int j = j0 + lane;
        uint32_t bal = __ballot_sync(FULL, j < p.N && sp[j] == p.hop);
        int c = __popc(bal);
        if (k < seen + c) {
          pos = j0 + (int)__fns(bal, 0, k - seen + 1);
          break;
        }
        seen += c;
      }
      slot = (int)__umulhi(r1, (uint32_t)p.max_nb);
    }
  }
  int j1 = -1;
  if (p.kind == QTX_SPIN_EXCHANGE) {
    j1 = p.nbr[pos * p.max_nb + slot];
    if (j1 < 0) j1 = pos;
  }
  const int8_t s0 = sp[pos], s1 = j1 >= 0 ? sp[j1] : 0;
  for (int j = lane; j < p.N; j += 32) {
    int8_t v = sp[j];
    if (p.kind == QTX_LOCAL_FLIP) {
      if (j == pos) v = -v;
    } else {
      if (j == pos) v = s1;
      if (j == j1) v = s0;
      if (j == pos && j1 == pos) v = s0;
    }
    np_[j] = v;
  }
  if (lane == 0) p.moved[chain] = (p.kind == QTX_LOCAL_FLIP) ? 1 : (s0 != s1);
}

struct AcceptParams {
  int8_t* spins;            // [ns, N] updated in place
  const int8_t* new_spins;
  const uint8_t* moved;
  double* mult; double* expo;            // current psi, updated in place
  const double* mult_new; const double* expo_new;
  int64_t ns; int N;
  double reweight;
  const double* inj_u;      // [ns] for this step (nullable -> Philox)
  uint32_t seed_lo, seed_hi;
  uint64_t step, chain0;
  int32_t* naccept;         // [ns] incremented (nullable)
  uint8_t* accept_log;      // [ns] for this step (nullable)
  const int32_t* rank;      // nullable: psi_new of chain c sits at index rank[c] (compacted batch of moved chains)
};

// metropolis.py:299-322: ratio = |psi'/psi|^n formed in the container then densified
// CPL: mult / mult_new are complex128 (interleaved re, im)
template <bool CPL>
__global__ void __launch_bounds__(256) accept_kernel(AcceptParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= p.ns) return;
  double u;
  if (p.inj_u) {
    u = p.inj_u[chain];
  } else {
    uint32_t r0 = (uint32_t)(p.chain0 + chain), r1 = (uint32_t)p.step, r2 = (uint32_t)(p.step >> 32), r3 = 0;
    philox4x32_10(r0, r1, r2, r3, p.seed_lo, p.seed_hi);
    u = (double)((((uint64_t)r2 << 32) | r3) >> 11) * 0x1.0p-53;
  }
  // unmoved proposals are never accepted (metropolis.py:314-316): their psi_new is not needed (and, with a
  // compacted batch, not computed)
  const bool moved = p.moved[chain] != 0;
  const int64_t cn = p.rank ? (moved ? (int64_t)p.rank[chain] : 0) : chain;
  const double e0 = p.expo[chain], e1 = p.expo_new[cn];
  double a0, a1, m1r, m1i = 0.0;
  if constexpr (CPL) {
    a0 = hypot(p.mult[2 * chain], p.mult[2 * chain + 1]);
    m1r = p.mult_new[2 * cn];
    m1i = p.mult_new[2 * cn + 1];
    a1 = hypot(m1r, m1i);
  } else {
    a0 = fabs(p.mult[chain]);
    m1r = p.mult_new[cn];
    a1 = fabs(m1r);
  }
  double rate = (a1 / a0) * exp(e1 - e0);
  rate = (p.reweight == 2.0) ? rate * rate : pow(rate, p.reweight);
  const bool zero_old = a0 * exp(e0) == 0.0;
  const bool acc = ((rate > 1.0 - u) || zero_old) && moved;
  if (acc) {
    int8_t* sp = p.spins + chain * p.N;
    const int8_t* np_ = p.new_spins + chain * p.N;
    for (int j = lane; j < p.N; j += 32) sp[j] = np_[j];
    if (lane == 0) {
      if constexpr (CPL) {
        p.mult[2 * chain] = m1r;
        p.mult[2 * chain + 1] = m1i;
      } else {
        p.mult[chain] = m1r;
      }
      p.expo[chain] = e1;
      if (p.naccept) p.naccept[chain] += 1;
    }
  }
  if (lane == 0 && p.accept_log) p.accept_log[chain] = acc ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// whole-network drivers
// ---------------------------------------------------------------------------------------------
struct NetShape {
  int nblocks, C, lx, ly, kh, kw, final_act;
  int N() const { return lx * ly; }
  int64_t w1_size(int i) const { return (int64_t)C * (i == 0 ? 1 : C) * kh * kw; }
  int64_t w2_size() const { return (int64_t)C * C * kh * kw; }
  int64_t nparams() const {
    int64_t n = 0;
    for (int i = 0; i < nblocks; ++i) n += w1_size(i) + C + w2_size() + (i == nblocks - 1 ? 0 : C);
    return n;
  }
};

static size_t resconv_ws_base(int dtype, int64_t ns, const NetShape& sh, bool grad) {
  size_t es = dtype == QTX_F64 ? 8 : 4;
  int64_t act = ns * sh.C * sh.N();
  int64_t elems = ns * sh.N();
  const int64_t wsz = (int64_t)sh.C * sh.C * sh.kh * sh.kw;
  if (grad) elems += (int64_t)(2 * sh.nblocks + 3) * act + 2 * wsz;
  else elems += 2 * act + wsz;
  return ((size_t)elems * es + 511) & ~(size_t)255;
}

template <typename T>
static int resconv_run(const NetShape& sh, const T* params, const int8_t* spins, int64_t ns, double* sig_out,
                       double* exp_out, void* out, int out_dtype, int64_t ld, void* ws, cudaStream_t st, int cpl = 0,
                       int64_t im_row_offset = 0, const long long* ns_dev = nullptr) {
  const int N = sh.N(), C = sh.C, nb = sh.nblocks;
  const bool grad = out != nullptr;
  const int64_t act = ns * C * N;
  T* base = reinterpret_cast<T*>(ws);
  T* x0 = base;                       // [ns, 1, N]
  T* X = x0 + ns * N;                 // grad: (nb+1) buffers X_1..X_nb (+ scratch); else 1 buffer
  T* Hs = X + (grad ? (int64_t)nb * act : act);   // grad: nb buffers; else 1
  T* scratch = Hs + (grad ? (int64_t)nb * act : act);  // grad: 3 gradient buffers + transposed weights; then repacked weights
  const int64_t wsz = (int64_t)C * C * sh.kh * sh.kw;
  T* wR = scratch + (grad ? 3 * act + wsz : 0);
  auto repack = [&](const T* w, int cout, int cin, int flip) -> int {
    weight_repack_kernel<T><<<64, 256, 0, st>>>(w, cout, cin, sh.kh, sh.kw, flip, wR);
    QTX_LAUNCH_CHECK();
    return QTX_OK;
  };
  bool tc_path = false;
  if constexpr (std::is_same<T, float>::value) tc_path = resconv_tc_supported(C, sh.lx, sh.ly, sh.kh, sh.kw);
  if (grad || !tc_path) {  // x0 feeds the CUDA-core first layer and the backward pass
    int64_t n = ns * N;
    unsigned g = (unsigned)((n + 255) / 256);
    if (g > 8u * num_sms()) g = 8u * num_sms();
    spins_to_act_kernel<T><<<g, 256, 0, st>>>(spins, n, x0);
    QTX_LAUNCH_CHECK();
  }
  // parameter offsets
  const T* pw1[64]; const T* pb1[64]; const T* pw2[64]; const T* pb2[64];
  int64_t col_w1[64], col_b1[64], col_w2[64], col_b2[64];
  QTX_REQUIRE(nb <= 64, QTX_ERR_UNSUPPORTED, "resconv: more than 64 blocks");
  {
    int64_t off = 0;
    for (int i = 0; i < nb; ++i) {
      pw1[i] = params + off; col_w1[i] = off; off += sh.w1_size(i);
      pb1[i] = params + off; col_b1[i] = off; off += C;
      pw2[i] = params + off; col_w2[i] = off; off += sh.w2_size();
      if (i == nb - 1) { pb2[i] = nullptr; col_b2[i] = -1; }
      else { pb2[i] = params + off; col_b2[i] = off; off += C; }
    }
  }
  // ---- forward (conv_nets.py:78-92) ----
  bool tc_done = false, tc_bwd = false;
  const T* tc_xfinal = nullptr;
  int tc_planes = 0;
  unsigned char* tc_ws = nullptr;
  size_t tc_ws_bytes = 0;
  if constexpr (std::is_same<T, float>::value) {
    if (resconv_tc_supported(C, sh.lx, sh.ly, sh.kh, sh.kw)) {
      // tensor-core tower (resconv_tc.cu); its workspace follows the regular one
      const size_t base_bytes = resconv_ws_base(QTX_F32, ns, sh, grad);
      tc_bwd = grad && resconv_tc_backward_supported(C, sh.lx, sh.ly, sh.kh, sh.kw);
      tc_ws = (unsigned char*)ws + base_bytes;
      tc_ws_bytes = resconv_tc_workspace(ns, nb, C, sh.lx, sh.ly, grad ? 1 : 0);
      int rc = resconv_tc_forward(nb, C, sh.lx, sh.ly, params, spins, ns, X, Hs, grad ? (tc_bwd ? 2 : 1) : 0, tc_ws,
                                  tc_ws_bytes, &tc_xfinal, &tc_planes, ns_dev, st);
      if (rc) return rc;
      tc_done = true;
    }
  }
  for (int i = 0; i < nb && !tc_done; ++i) {
    const T* xin = (i == 0) ? x0 : (grad ? X + (int64_t)(i - 1) * act : X);
    T* h = grad ? Hs + (int64_t)i * act : Hs;
    T* xout = grad ? X + (int64_t)i * act : X;
    ConvParams<T> p{};
    p.ns = ns; p.lx = sh.lx; p.ly = sh.ly; p.kh = sh.kh; p.kw = sh.kw;
    p.in = xin; p.cin = (i == 0) ? 1 : C; p.cout = C; p.w = pw1[i]; p.bias = pb1[i];
    p.alpha = (i == 0) ? (T)(1.0 / sqrt(2.0)) : (T)(1.0 / sqrt((double)(i + 1)));
    p.in_act = (i == 0) ? 0 : 1;
    p.res = nullptr; p.res_ch = 0; p.mul = nullptr; p.out = h;
    int rc = repack(pw1[i], C, p.cin, 0);
    if (rc) return rc;
    p.w_ctco = wR;
    rc = launch_conv<T>(p, st);
    if (rc) return rc;
    ConvParams<T> q{};
    q.ns = ns; q.lx = sh.lx; q.ly = sh.ly; q.kh = sh.kh; q.kw = sh.kw;
    q.in = h; q.cin = C; q.cout = C; q.w = pw2[i]; q.bias = pb2[i]; q.alpha = 1; q.in_act = 1;
    q.res = xin; q.res_ch = (i == 0) ? 1 : C; q.mul = nullptr; q.out = xout;
    if ((rc = repack(pw2[i], C, C, 0))) return rc;
    q.w_ctco = wR;
    rc = launch_conv<T>(q, st);
    if (rc) return rc;
  }
  const T* xlast = tc_done ? tc_xfinal : (grad ? X + (int64_t)(nb - 1) * act : X);
  T* dA = grad ? scratch : nullptr;           // gradient w.r.t. the current block output
  T* dB = grad ? scratch + act : nullptr;
  T* dC = grad ? scratch + 2 * act : nullptr; // complex output: seed of the imaginary-part pass
  T* wT = grad ? scratch + 3 * act : nullptr;
  if (cpl)
    resconv_final_cplx_kernel<T><<<(unsigned)ns, 256, 0, st>>>(xlast, ns, C, N, (T)(1.0 / sqrt((double)(nb + 1))),
                                                               sh.final_act, (double2*)sig_out, exp_out, dA, dC, tc_planes, ns_dev);
  else if (std::is_same<T, float>::value && tc_planes > 0 && !dA && !getenv("QTX_FINAL_TWO_PASS"))
    resconv_final_planar_kernel<<<(unsigned)ns, 256, 0, st>>>((const float*)xlast, ns, C, N, (float)(1.0 / sqrt((double)(nb + 1))),
                                                              sh.final_act, sig_out, exp_out, tc_planes, ns_dev);
  else
    resconv_final_kernel<T><<<(unsigned)ns, 256, 0, st>>>(xlast, ns, C, N, (T)(1.0 / sqrt((double)(nb + 1))),
                                                          sh.final_act, sig_out, exp_out, dA, tc_planes, ns_dev);
  QTX_LAUNCH_CHECK();
  if (!grad) return QTX_OK;

  // ---- backward: per-sample parameter gradients straight into the Jacobian rows ----
  const int taps = sh.kh * sh.kw;
  auto wgrad = [&](const T* delta, const T* a, int cin, T alpha, int in_act, int64_t col0) -> int {
    size_t smem = ((size_t)32 * N + (size_t)kCK * (sh.lx + sh.kh - 1) * (sh.ly + sh.kw - 1)) * sizeof(T);
    dim3 grid((unsigned)ns, (unsigned)((C + 31) / 32));
    if (out_dtype == QTX_F64) {
      auto k = conv_wgrad_kernel<T, double>;
      if (smem > 48 * 1024) QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, 192, smem, st>>>(delta, a, cin, C, sh.lx, sh.ly, sh.kh, sh.kw, alpha, in_act, (double*)out, ld, col0);
    } else {
      auto k = conv_wgrad_kernel<T, float>;
      if (smem > 48 * 1024) QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, 192, smem, st>>>(delta, a, cin, C, sh.lx, sh.ly, sh.kh, sh.kw, alpha, in_act, (float*)out, ld, col0);
    }
    QTX_LAUNCH_CHECK();
    return QTX_OK;
  };
  auto bgrad = [&](const T* delta, int64_t col0) -> int {
    unsigned g = (unsigned)((ns * C + 7) / 8);
    if (out_dtype == QTX_F64) conv_bgrad_kernel<T, double><<<g, 256, 0, st>>>(delta, ns, C, N, (double*)out, ld, col0);
    else conv_bgrad_kernel<T, float><<<g, 256, 0, st>>>(delta, ns, C, N, (float*)out, ld, col0);
    QTX_LAUNCH_CHECK();
    return QTX_OK;
  };
  // complex output: a second backward pass seeded with d Im(log psi) writes rows [im_row_offset, im_row_offset + ns)
  for (int pass = 0; pass < (cpl ? 2 : 1); ++pass) {
  if (pass == 1) {
    const size_t esz = out_dtype == QTX_F64 ? 8 : 4;
    out = (unsigned char*)out + (size_t)im_row_offset * ld * esz;
  }
  T* dX = pass == 0 ? dA : dC;  // gradient w.r.t. X_{i+1}
  T* tmp = dB;
  if constexpr (std::is_same<T, float>::value) {
    if (tc_bwd) {
      // tensor-core Jacobian (resconv_tc.cu): backward-data tower + per-sample weight gradients of the C x C
      // convolutions; biases and the one-channel first convolution from the raw output gradients it leaves behind
      const float* rg[2 * 64 + 1];
      int rc = resconv_tc_backward(nb, C, sh.lx, sh.ly, params, ns, X, Hs, dX, out, out_dtype == QTX_F64 ? 1 : 0, ld,
                                   tc_ws, tc_ws_bytes, rg, st);
      if (rc) return rc;
      for (int i = nb - 1; i >= 0; --i) {
        const int k = 2 * (nb - 1 - i);
        if (col_b2[i] >= 0 && (rc = bgrad(rg[k], col_b2[i]))) return rc;
        if ((rc = bgrad(rg[k + 1], col_b1[i]))) return rc;
      }
      if ((rc = wgrad(rg[2 * nb - 1], x0, 1, (T)(1.0 / sqrt(2.0)), 0, col_w1[0]))) return rc;
      const char* e = getenv("QTX_TC_WGRAD");
      if (e && e[0] == '0') {  // dev knob: weight gradients on the CUDA cores from the tower's raw gradients
        for (int i = nb - 1; i >= 0; --i) {
          const int k = 2 * (nb - 1 - i);
          if ((rc = wgrad(rg[k], Hs + (int64_t)i * act, C, (T)1, 1, col_w2[i]))) return rc;
          if (i > 0 && (rc = wgrad(rg[k + 1], X + (int64_t)(i - 1) * act, C, (T)(1.0 / sqrt((double)(i + 1))), 1, col_w1[i])))
            return rc;
        }
      }
      continue;
    }
  }
  for (int i = nb - 1; i >= 0; --i) {
    const T* xin = (i == 0) ? x0 : X + (int64_t)(i - 1) * act;
    const T* h = Hs + (int64_t)i * act;
    int rc;
    // conv2: y = conv2(gelu(h)) + b2
    if ((rc = wgrad(dX, h, C, (T)1, 1, col_w2[i]))) return rc;
    if (col_b2[i] >= 0 && (rc = bgrad(dX, col_b2[i]))) return rc;
    // dh = conv2^T(dX) * gelu'(h)
    weight_transpose_flip_kernel<T><<<64, 256, 0, st>>>(pw2[i], C, C, sh.kh, sh.kw, wT);
    QTX_LAUNCH_CHECK();
    ConvParams<T> p{};
    p.ns = ns; p.lx = sh.lx; p.ly = sh.ly; p.kh = sh.kh; p.kw = sh.kw;
    p.in = dX; p.cin = C; p.cout = C; p.w = wT; p.bias = nullptr; p.alpha = 1; p.in_act = 0;
    p.res = nullptr; p.res_ch = 0; p.mul = h; p.mul_alpha = 1; p.mul_scale = 1; p.out = tmp;
    if ((rc = repack(pw2[i], C, C, 1))) return rc;
    p.w_ctco = wR;
    if ((rc = launch_conv<T>(p, st))) return rc;
    // conv1: h = conv1(a1) + b1
    const T alpha1 = (i == 0) ? (T)(1.0 / sqrt(2.0)) : (T)(1.0 / sqrt((double)(i + 1)));
    if ((rc = wgrad(tmp, xin, (i == 0) ? 1 : C, alpha1, (i == 0) ? 0 : 1, col_w1[i]))) return rc;
    if ((rc = bgrad(tmp, col_b1[i]))) return rc;
    if (i > 0) {
      // dX_in = conv1^T(dh) * gelu'(x/sqrt(i+1)) / sqrt(i+1) + dX   (residual)
      weight_transpose_flip_kernel<T><<<64, 256, 0, st>>>(pw1[i], C, C, sh.kh, sh.kw, wT);
      QTX_LAUNCH_CHECK();
      ConvParams<T> q{};
      q.ns = ns; q.lx = sh.lx; q.ly = sh.ly; q.kh = sh.kh; q.kw = sh.kw;
      q.in = tmp; q.cin = C; q.cout = C; q.w = wT; q.bias = nullptr; q.alpha = 1; q.in_act = 0;
      q.res = dX; q.res_ch = C; q.mul = xin; q.mul_alpha = alpha1; q.mul_scale = alpha1;
      // in-place on dX is safe: each thread reads res at exactly the index it writes
      q.out = dX;
      if ((rc = repack(pw1[i], C, C, 1))) return rc;
      q.w_ctco = wR;
      if ((rc = launch_conv<T>(q, st))) return rc;
    }
  }
  }
  (void)taps;
  return QTX_OK;
}

static size_t resconv_ws(int dtype, int64_t ns, const NetShape& sh, bool grad) {
  size_t b = resconv_ws_base(dtype, ns, sh, grad);
  if (dtype == QTX_F32 && resconv_tc_supported(sh.C, sh.lx, sh.ly, sh.kh, sh.kw))
    b += resconv_tc_workspace(ns, sh.nblocks, sh.C, sh.lx, sh.ly, grad ? 1 : 0);
  return b;
}

}  // namespace qtx

using namespace qtx;

static bool shape_ok(int nblocks, int C, int lx, int ly, int kh, int kw, int final_act) {
  return nblocks >= 1 && C >= 1 && lx >= 1 && ly >= 1 && kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1) && kh <= lx + 1 &&
         (final_act == 0 || final_act == 1);
}

extern "C" int qtx_resconv_tc_available(int model_dtype, int channels, int lx, int ly, int kh, int kw) {
  return (model_dtype == QTX_F32 && resconv_tc_supported(channels, lx, ly, kh, kw)) ? 1 : 0;
}

extern "C" int qtx_resconv_tc_backward_available(int model_dtype, int channels, int lx, int ly, int kh, int kw) {
  return (model_dtype == QTX_F32 && resconv_tc_backward_supported(channels, lx, ly, kh, kw)) ? 1 : 0;
}

extern "C" int64_t qtx_resconv_nparams(int nblocks, int channels, int lx, int ly, int kh, int kw) {
  NetShape sh{nblocks, channels, lx, ly, kh, kw, 0};
  return sh.nparams();
}

extern "C" size_t qtx_resconv_workspace_size(int model_dtype, int64_t ns, int nblocks, int channels, int lx, int ly,
                                             int kh, int kw, int need_grad) {
  NetShape sh{nblocks, channels, lx, ly, kh, kw, 0};
  return resconv_ws(model_dtype, ns, sh, need_grad != 0);
}

extern "C" int qtx_resconv_forward(int model_dtype, const void* params, int nblocks, int channels, int lx, int ly,
                                   int kh, int kw, int final_act, const int8_t* spins, int64_t ns,
                                   double* significand_out, double* exponent_out, void* workspace,
                                   size_t workspace_bytes, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(params && spins && significand_out && exponent_out && workspace, QTX_ERR_INVALID,
              "qtx_resconv_forward: bad argument");
  QTX_REQUIRE(shape_ok(nblocks, channels, lx, ly, kh, kw, final_act), QTX_ERR_INVALID,
              "qtx_resconv_forward: bad network shape");
  NetShape sh{nblocks, channels, lx, ly, kh, kw, final_act};
  QTX_REQUIRE(workspace_bytes >= resconv_ws(model_dtype, ns, sh, false), QTX_ERR_INVALID,
              "qtx_resconv_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32)
    return resconv_run<float>(sh, (const float*)params, spins, ns, significand_out, exponent_out, nullptr, 0, 0,
                              workspace, st);
  if (model_dtype == QTX_F64)
    return resconv_run<double>(sh, (const double*)params, spins, ns, significand_out, exponent_out, nullptr, 0, 0,
                               workspace, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_resconv_forward: bad dtype %d", model_dtype);
}

extern "C" int qtx_resconv_jacobian(int model_dtype, const void* params, int nblocks, int channels, int lx, int ly,
                                    int kh, int kw, int final_act, const int8_t* spins, int64_t ns, int out_dtype,
                                    void* out, int64_t ld, double* significand_out, double* exponent_out,
                                    void* workspace, size_t workspace_bytes, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(params && spins && out && workspace, QTX_ERR_INVALID, "qtx_resconv_jacobian: bad argument");
  QTX_REQUIRE(shape_ok(nblocks, channels, lx, ly, kh, kw, final_act), QTX_ERR_INVALID,
              "qtx_resconv_jacobian: bad network shape");
  NetShape sh{nblocks, channels, lx, ly, kh, kw, final_act};
  QTX_REQUIRE(ld >= sh.nparams(), QTX_ERR_INVALID, "qtx_resconv_jacobian: ld smaller than the parameter count");
  QTX_REQUIRE(out_dtype == QTX_F32 || out_dtype == QTX_F64, QTX_ERR_INVALID, "qtx_resconv_jacobian: bad out dtype");
  QTX_REQUIRE(workspace_bytes >= resconv_ws(model_dtype, ns, sh, true), QTX_ERR_INVALID,
              "qtx_resconv_jacobian: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32)
    return resconv_run<float>(sh, (const float*)params, spins, ns, significand_out, exponent_out, out, out_dtype, ld,
                              workspace, st);
  if (model_dtype == QTX_F64)
    return resconv_run<double>(sh, (const double*)params, spins, ns, significand_out, exponent_out, out, out_dtype,
                               ld, workspace, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_resconv_jacobian: bad dtype %d", model_dtype);
}

// ---- batches of MOVED proposals only (the reference evaluates psi of every proposal, also of the no-ops that can
// never be accepted, metropolis.py:262-275,314-316): rank[c] = index of chain c among the moved chains (or -1),
// their proposed configurations gathered contiguously, the count left on the device ----
__global__ void __launch_bounds__(1024) compact_moved_kernel(const uint8_t* __restrict__ moved, int64_t ns,
                                                             int32_t* __restrict__ rank, long long* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < ns; base += blockDim.x) {
    const int64_t c = base + tid;
    const int f = (c < ns && moved[c]) ? 1 : 0;
    int incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    const int carry = carry_s;
    if (c < ns) rank[c] = f ? carry + woff + incl - 1 : -1;
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = carry + woff + incl;
    __syncthreads();
  }
  if (tid == 0) *count = carry_s;
}

__global__ void __launch_bounds__(256) gather_moved_kernel(const int8_t* __restrict__ new_spins, const int32_t* __restrict__ rank,
                                                           int64_t ns, int N, int8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= ns) return;
  const int r = rank[chain];
  if (r < 0) return;
  for (int j = lane; j < N; j += 32) out[(int64_t)r * N + j] = new_spins[chain * N + j];
}

extern "C" int qtx_compact_moved(const uint8_t* moved, const int8_t* new_spins, int64_t ns, int N, int32_t* rank_out,
                                 int8_t* compact_spins_out, int64_t* count_out, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(moved && new_spins && rank_out && compact_spins_out && count_out && N > 0, QTX_ERR_INVALID,
              "qtx_compact_moved: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  compact_moved_kernel<<<1, 1024, 0, st>>>(moved, ns, rank_out, (long long*)count_out);
  QTX_LAUNCH_CHECK();
  gather_moved_kernel<<<(unsigned)((ns + 7) / 8), 256, 0, st>>>(new_spins, rank_out, ns, N, compact_spins_out);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_metropolis_accept_compact(int8_t* spins, const int8_t* new_spins, const uint8_t* moved,
                                             const int32_t* rank, int64_t ns, int N, double* mult, double* expo,
                                             const double* mult_new, const double* expo_new, int mult_complex,
                                             double reweight, const double* inj_u, uint64_t seed, uint64_t step,
                                             uint64_t chain0, int32_t* naccept, uint8_t* accept_log,
                                             qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(spins && new_spins && moved && rank && mult && expo && mult_new && expo_new && N > 0, QTX_ERR_INVALID,
              "qtx_metropolis_accept_compact: bad argument");
  AcceptParams p;
  p.spins = spins; p.new_spins = new_spins; p.moved = moved; p.mult = mult; p.expo = expo; p.mult_new = mult_new;
  p.expo_new = expo_new; p.ns = ns; p.N = N; p.reweight = reweight; p.inj_u = inj_u;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.step = step; p.chain0 = chain0;
  p.naccept = naccept; p.accept_log = accept_log; p.rank = rank;
  if (mult_complex) accept_kernel<true><<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  else accept_kernel<false><<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

// Forward of the first *ns_dev (<= ns_max) samples, the count living on the device (no host synchronisation between
// the proposal and the forward).  Float32 3x3 towers on the tensor-core path only; QTX_ERR_UNSUPPORTED otherwise.
extern "C" int qtx_resconv_forward_n(int model_dtype, const void* params, int nblocks, int channels, int lx, int ly,
                                     int kh, int kw, int final_act, int out_complex, const int8_t* spins, int64_t ns_max,
                                     const int64_t* ns_dev, double* significand_out, double* exponent_out,
                                     void* workspace, size_t workspace_bytes, qtx_stream_t stream) {
  if (ns_max == 0) return QTX_OK;
  QTX_REQUIRE(params && spins && ns_dev && significand_out && exponent_out && workspace, QTX_ERR_INVALID,
              "qtx_resconv_forward_n: bad argument");
  QTX_REQUIRE(shape_ok(nblocks, channels, lx, ly, kh, kw, final_act) && (!out_complex || channels % 2 == 0),
              QTX_ERR_INVALID, "qtx_resconv_forward_n: bad network shape");
  QTX_REQUIRE(model_dtype == QTX_F32 && resconv_tc_supported(channels, lx, ly, kh, kw), QTX_ERR_UNSUPPORTED,
              "qtx_resconv_forward_n: needs the float32 tensor-core tower");
  NetShape sh{nblocks, channels, lx, ly, kh, kw, final_act};
  QTX_REQUIRE(workspace_bytes >= resconv_ws(model_dtype, ns_max, sh, false), QTX_ERR_INVALID,
              "qtx_resconv_forward_n: workspace too small");
  return resconv_run<float>(sh, (const float*)params, spins, ns_max, significand_out, exponent_out, nullptr, 0, 0,
                            workspace, (cudaStream_t)stream, out_complex ? 1 : 0, 0, (const long long*)ns_dev);
}

extern "C" int qtx_resconv_forward_cplx(int model_dtype, const void* params, int nblocks, int channels, int lx, int ly,
                                        int kh, int kw, int final_act, const int8_t* spins, int64_t ns,
                                        double* significand_c128_out, double* exponent_out, void* workspace,
                                        size_t workspace_bytes, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(params && spins && significand_c128_out && exponent_out && workspace, QTX_ERR_INVALID,
              "qtx_resconv_forward_cplx: bad argument");
  QTX_REQUIRE(shape_ok(nblocks, channels, lx, ly, kh, kw, final_act) && channels % 2 == 0, QTX_ERR_INVALID,
              "qtx_resconv_forward_cplx: bad network shape (channels must be even)");
  NetShape sh{nblocks, channels, lx, ly, kh, kw, final_act};
  QTX_REQUIRE(workspace_bytes >= resconv_ws(model_dtype, ns, sh, false), QTX_ERR_INVALID,
              "qtx_resconv_forward_cplx: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32)
    return resconv_run<float>(sh, (const float*)params, spins, ns, significand_c128_out, exponent_out, nullptr, 0, 0,
                              workspace, st, 1, 0);
  if (model_dtype == QTX_F64)
    return resconv_run<double>(sh, (const double*)params, spins, ns, significand_c128_out, exponent_out, nullptr, 0, 0,
                               workspace, st, 1, 0);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_resconv_forward_cplx: bad dtype %d", model_dtype);
}

extern "C" int qtx_resconv_jacobian_cplx(int model_dtype, const void* params, int nblocks, int channels, int lx, int ly,
                                         int kh, int kw, int final_act, const int8_t* spins, int64_t ns, int out_dtype,
                                         void* out, int64_t ld, int64_t im_row_offset, double* significand_c128_out,
                                         double* exponent_out, void* workspace, size_t workspace_bytes,
                                         qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(params && spins && out && workspace, QTX_ERR_INVALID, "qtx_resconv_jacobian_cplx: bad argument");
  QTX_REQUIRE(shape_ok(nblocks, channels, lx, ly, kh, kw, final_act) && channels % 2 == 0, QTX_ERR_INVALID,
              "qtx_resconv_jacobian_cplx: bad network shape (channels must be even)");
  NetShape sh{nblocks, channels, lx, ly, kh, kw, final_act};
  QTX_REQUIRE(ld >= sh.nparams(), QTX_ERR_INVALID, "qtx_resconv_jacobian_cplx: ld smaller than the parameter count");
  QTX_REQUIRE(im_row_offset >= ns, QTX_ERR_INVALID, "qtx_resconv_jacobian_cplx: the imaginary block overlaps the real one");
  QTX_REQUIRE(out_dtype == QTX_F32 || out_dtype == QTX_F64, QTX_ERR_INVALID, "qtx_resconv_jacobian_cplx: bad out dtype");
  QTX_REQUIRE(workspace_bytes >= resconv_ws(model_dtype, ns, sh, true), QTX_ERR_INVALID,
              "qtx_resconv_jacobian_cplx: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32)
    return resconv_run<float>(sh, (const float*)params, spins, ns, significand_c128_out, exponent_out, out, out_dtype, ld,
                              workspace, st, 1, im_row_offset);
  if (model_dtype == QTX_F64)
    return resconv_run<double>(sh, (const double*)params, spins, ns, significand_c128_out, exponent_out, out, out_dtype,
                               ld, workspace, st, 1, im_row_offset);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_resconv_jacobian_cplx: bad dtype %d", model_dtype);
}

extern "C" int qtx_metropolis_propose(int kind, const int8_t* spins, int64_t ns, int N, const int32_t* nbr_table,
                                      int max_nb, int hop, const int32_t* inj_pos, const int32_t* inj_slot,
                                      uint64_t seed, uint64_t step, uint64_t chain0, int8_t* new_spins,
                                      uint8_t* moved, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(spins && new_spins && moved && N > 0, QTX_ERR_INVALID, "qtx_metropolis_propose: bad argument");
  QTX_REQUIRE(kind == QTX_LOCAL_FLIP || (kind == QTX_SPIN_EXCHANGE && nbr_table && max_nb > 0), QTX_ERR_INVALID,
              "qtx_metropolis_propose: bad kind / neighbour table");
  ProposeParams p;
  p.spins = spins; p.new_spins = new_spins; p.moved = moved; p.ns = ns; p.N = N; p.kind = kind; p.max_nb = max_nb;
  p.hop = hop; p.nbr = nbr_table; p.inj_pos = inj_pos; p.inj_slot = inj_slot;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.step = step; p.chain0 = chain0;
  propose_kernel<<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_metropolis_accept(int8_t* spins, const int8_t* new_spins, const uint8_t* moved, int64_t ns, int N,
                                     double* mult, double* expo, const double* mult_new, const double* expo_new,
                                     double reweight, const double* inj_u, uint64_t seed, uint64_t step,
                                     uint64_t chain0, int32_t* naccept, uint8_t* accept_log, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(spins && new_spins && moved && mult && expo && mult_new && expo_new && N > 0, QTX_ERR_INVALID,
              "qtx_metropolis_accept: bad argument");
  AcceptParams p;
  p.spins = spins; p.new_spins = new_spins; p.moved = moved; p.mult = mult; p.expo = expo; p.mult_new = mult_new;
  p.expo_new = expo_new; p.ns = ns; p.N = N; p.reweight = reweight; p.inj_u = inj_u;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.step = step; p.chain0 = chain0;
  p.naccept = naccept; p.accept_log = accept_log; p.rank = nullptr;
  accept_kernel<false><<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_metropolis_accept_cplx(int8_t* spins, const int8_t* new_spins, const uint8_t* moved, int64_t ns, int N,
                                          double* mult, double* expo, const double* mult_new, const double* expo_new,
                                          double reweight, const double* inj_u, uint64_t seed, uint64_t step,
                                          uint64_t chain0, int32_t* naccept, uint8_t* accept_log, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(spins && new_spins && moved && mult && expo && mult_new && expo_new && N > 0, QTX_ERR_INVALID,
              "qtx_metropolis_accept_cplx: bad argument");
  AcceptParams p;
  p.spins = spins; p.new_spins = new_spins; p.moved = moved; p.mult = mult; p.expo = expo; p.mult_new = mult_new;
  p.expo_new = expo_new; p.ns = ns; p.N = N; p.reweight = reweight; p.inj_u = inj_u;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.step = step; p.chain0 = chain0;
  p.naccept = naccept; p.accept_log = accept_log; p.rank = nullptr;
  accept_kernel<true><<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

// psi *= exp(i * sum_j kernel[j] * s_j)   (quantax/nn/sign.py:8-43 with output="phase"; the dot product is float32
// like the reference's kernel.astype(float32)); warp per sample, mult is complex128
__global__ void __launch_bounds__(256) sign_phase_kernel(const float* __restrict__ kernel, const int8_t* __restrict__ spins,
                                                         int64_t ns, int N, double2* __restrict__ mult) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= ns) return;
  float acc = 0.f;
  for (int j = lane; j < N; j += 32) acc += kernel[j] * (float)spins[s * N + j];
  acc = warp_sum(acc);
  if (lane == 0) {
    float sn, cs;
    sincosf(acc, &sn, &cs);
    const double2 m = mult[s];
    mult[s] = make_double2(m.x * (double)cs - m.y * (double)sn, m.x * (double)sn + m.y * (double)cs);
  }
}

extern "C" int qtx_apply_sign_phase(const float* kernel, const int8_t* spins, int64_t ns, int N, double* mult_c128,
                                    qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(kernel && spins && mult_c128 && N > 0, QTX_ERR_INVALID, "qtx_apply_sign_phase: bad argument");
  sign_phase_kernel<<<(unsigned)((ns + 7) / 8), 256, 0, (cudaStream_t)stream>>>(kernel, spins, ns, N, (double2*)mult_c128);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}
