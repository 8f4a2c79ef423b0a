// ResConv forward on the 5th-generation tensor cores (tcgen05 / TMEM), float32 models.
//
// quantax/model/conv_nets.py:78-92 — every 3x3 circular convolution of the residual tower is an
// implicit GEMM  D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, cin, tap]  with
// M = pixels, N = cout, K = 9 * cin.  float32 accuracy (parity bar 1e-5) is kept with a two-term
// binary16 split of both operands (x = hi + lo, |lo| <= 2^-11 |hi|) and three kind::f16 products
// hi*hi + lo*hi + hi*lo accumulated in float32 in TMEM: the neglected lo*lo term and the rounding of
// lo are ~2^-22 relative, below float32 GEMM rounding.  Operands are pre-scaled by powers of two
// (activations x4, weights x256) so that `lo` stays a normal binary16 number for all values that
// matter; the epilogue removes the exact factor 2^-10.  Range: |gelu output| < 16376.
//
// Layout ("raster"): the gelu'ed, split input of a convolution is stored per sample as a padded
// raster of 16-byte slots (8 binary16 channels per slot), circular halo included, one plane per
// 8-channel group:  act[kstep][hi|lo][plane-in-kstep][slot][8].  In the no-swizzle K-major UMMA
// layout a core matrix is 8 rows x 16 B with the rows 16 B apart, so 8 consecutive slots of one plane
// ARE a core matrix, and the operand tile of tap (dy, dx) is the same shared-memory tile addressed
// (dy * row_pitch + dx) slots further: the nine taps re-read one staged tile, nothing is duplicated
// in shared memory and no im2col is materialised.  Two rasters are used:
//   SEG    (W % 8 == 0):  rows are cut into 8-pixel segments, each stored with its own left/right
//          halo slot (10 slots); M rows = interior pixels only, stride-byte-offset = 10 slots
//          -> no wasted MMA rows (16x16: one sample = two 128-row tiles);
//   RASTER (any W):       plain (W+2)-pitch raster, M rows = consecutive slots from the first to the
//          last interior pixel, halo columns compute garbage rows that the epilogue masks.
//
// Tensor-core accumulation truncates (round toward zero) at every accumulate, a systematic bias of
// ~0.5 ulp per step; the hi*hi products go to one TMEM accumulator (54 steps at C = 88) and the two
// small cross products to a second one, whose truncation is 2^-11 times smaller; the epilogue adds
// the two in round-to-nearest float32.
//
// One persistent kernel runs ALL tensor-core layers of the tower: a CTA owns its samples through the
// whole depth (a layer's output of a sample depends only on that sample), alternates between two
// work items (the epilogue of one overlaps the MMAs of the other: the accumulators are drained into
// registers first and TMEM is released at once) and updates the operand buffer in place, so
// activations stay L2-resident and only weights stream.
//   warp 0     bulk-copy producer (cp.async.bulk, mbarrier complete_tx): activation k-step tiles and
//              weight (k-step, kernel-row) tiles, two rings
//   warp 1     MMA issuer (one thread), TMEM owner
//   warps 2-13 epilogue: TMEM -> registers, bias / residual, raw float32 output (next residual, and
//              the saved activations of the backward pass), gelu, split, operand store with halo copies
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "gram_tc_common.cuh"

namespace qtx {

constexpr int kTcThreads = 448;
constexpr int kEpiWarps = 12;
constexpr int kColGroups = kEpiWarps / 4;
constexpr int kTcMaxLayers = 32;
constexpr float kActScale = 4.0f;
constexpr float kWScale = 256.0f;
constexpr float kOutScale = 1.0f / (kActScale * kWScale);

struct TcGeom {
  int H, W, mode;  // mode 1 = SEG, 0 = RASTER
  int nseg, RP, Ps, SB, TS;
  int spi, tps;    // samples per item, tiles per sample (spi * tps == 2; == 4 for CTA pairs)
  int pair, TSh;   // CTA-pair kernel: each CTA stages TSh slots per pair-tile
  int64_t ns;      // samples
  int64_t nitems;
  int64_t slots;   // slots per plane (all samples + tail guard)
};

struct TcLayer {
  int64_t wblob_off;       // in halfs
  const float* bias;       // [Np] zero-padded copy in the workspace (zeros when the conv has no bias)
  const float* res;        // raw residual or null
  const int8_t* res_spin;  // block 0 residual: the spins, broadcast over channels, or null
  float* raw_out;          // raw output or null
  float out_alpha;         // operand out = split(kActScale * gelu(out_alpha * v))
  int write_act;
  int planar;              // raw layout: 0 = [ns, C, N], 1 = [ns, Np/8, N, 8] (forward-only residual stream)
  // operand buffers (units of one raster set [KS][hi|lo][2][slots]): the forward-only tower works in place (0, 0),
  // the towers of the Jacobian keep every layer's operand (CTA-pair kernel only)
  int in_buf, out_buf;
  // mode 1 = backward-data layer (variational.py:429-491 through conv_nets.py:78-92):
  //   v = conv^T(g) * mul_alpha * gelu'(mul_alpha * mul) + res,  operand out = split(sig_out * v)
  // with per-sample power-of-two scales so that the binary16 pair keeps float32 accuracy at any gradient magnitude
  int mode;
  const float* mul;        // raw pre-activation [ns, C, N]
  float mul_alpha;
  const float* sig_in;     // [ns] scale of the incoming gradient operand
  const float* max_in;     // [ns] max |g| of the incoming gradient
  const float* max_res;    // [ns] max |res| or null
  const float* wnorm;      // [1] max_c sum_{o,tap} |w[o,c,tap]|: |conv^T(g)| <= wnorm * max|g|
  float* sig_out;          // [ns]
  float* max_out;          // [ns] float bits, atomicMax (zeroed by the host)
};

struct TcNetParams {
  __half* act;
  const __half* wblob;
  TcGeom g;
  int C, Np, KS;
  int layer0, layer1;  // layers [layer0, layer1) are run by this launch
  int act_stages, w_stages;
  int64_t buf_u4;          // 16-byte units per operand buffer
  int precise;             // accurate gelu / gelu' (towers of the Jacobian)
  __half* act_z;           // backward-data tower on the RASTER layout: second operand set written WITHOUT halo copies
                           // (the K axis of the weight-gradient GEMM crosses the halo columns), or null
  float out_scale;         // kOutScale times the expected-value correction of the truncating accumulator (tc_trunc_comp)
  int epi_generic;         // dev knob QTX_TC_EPI_GENERIC=1: the generic epilogue for every layer
  int pair_k;              // CTA-pair kernel, C % 16 in 1..8: the half-filled last K step pairs two taps per MMA
  int keep_pad;            // dev knob QTX_TC_KEEP_PAD=1: the forward-only epilogues also process the padded channel group
  unsigned long long* dbg;  // optional [grid][16] cycle counters (QTX_TC_DEBUG=1)
  const long long* ns_dev;  // optional device-side sample count (<= g.ns): batches whose size is decided on the device
  TcLayer layer[kTcMaxLayers];
};

// wait on an mbarrier and charge the waiting time to a counter (profiling aid; `t` lives in a register)
#define QTX_TIMED_WAIT(counter, bar, parity)        \
  do {                                              \
    const long long _t0 = clock64();                \
    mbar_wait((bar), (parity));                     \
    (counter) += (unsigned long long)(clock64() - _t0); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as (low word, high word): the high word (SBO, version) is loop-invariant and the
// low word (address, LBO) is advanced with one 32-bit add per operand
__device__ __forceinline__ void umma_f16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp; ptxas keeps warp-uniform operands of the elected region in uniform registers
// (UTCHMMA takes uniform-register operands: without this every MMA pays a vector->uniform election loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// no wait: the caller issues tcgen05.wait::ld once after a batch of loads
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// K-major, no swizzle: core matrix = 8 rows x 16 B (rows 16 B apart); LBO = distance between the two
// core matrices of one K step, SBO = distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}

// kActScale * gelu(alpha * v) with the constants folded:
//   gelu(x) = x * sigmoid(2 sqrt(2/pi) (x + 0.044715 x^3)) = x / (1 + 2^(x (a + b x^2)))
struct GeluConst {
  float a, b, c;  // exponent polynomial in v, output factor
  float alpha;
  int precise;    // accurate exp and division (the forward pass that feeds the Jacobian)
};
__host__ __device__ inline GeluConst gelu_const(float alpha, int precise = 0) {
  GeluConst k;
  k.alpha = alpha;
  k.precise = precise;
  const float l2e = 1.4426950408889634f, s = -1.5957691216057308f;
  k.a = s * l2e * alpha;
  k.b = s * l2e * 0.044715f * alpha * alpha * alpha;
  k.c = kActScale * alpha;
  return k;
}
__device__ __forceinline__ float gelu_scaled(float v, const GeluConst& k) {
  if (k.precise) {
    const float x = k.alpha * v;
    const float u = 1.5957691216057308f * (x + 0.044715f * x * x * x);
    return kActScale * (x / (1.0f + expf(-u)));
  }
  const float t = v * fmaf(v * v, k.b, k.a);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return (k.c * v) * r;
}

// d/dx gelu(x) (tanh form): s + x s (1 - s) 2 sqrt(2/pi) (1 + 3 * 0.044715 x^2),  s = sigmoid(2 sqrt(2/pi) (x + 0.044715 x^3))
__device__ __forceinline__ float gelu_grad_fast(float x, int precise = 0) {
  const float l2e = 1.4426950408889634f, c = 1.5957691216057308f;
  const float x2 = x * x;
  if (precise) {
    const float sg = 1.0f / (1.0f + expf(-c * x * fmaf(0.044715f, x2, 1.0f)));
    return fmaf(x * sg * (1.0f - sg), c * fmaf(3.0f * 0.044715f, x2, 1.0f), sg);
  }
  const float t = -c * l2e * x * fmaf(0.044715f, x2, 1.0f);
  float e, sg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + e));
  return fmaf(x * sg * (1.0f - sg), c * fmaf(3.0f * 0.044715f, x2, 1.0f), sg);
}
// power of two that maps |v| <= bound below 2^15 (binary16 overflows at 65504)
__device__ __forceinline__ float grad_scale(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.0f;
  int e;
  (void)frexpf(bound, &e);  // bound = m 2^e, m in [0.5, 1)
  int k = 15 - e;
  k = k < -100 ? -100 : (k > 100 ? 100 : k);
  return __int_as_float((k + 127) << 23);
}

// x = hi + lo in binary16, two values at a time
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// the operand slots (relative to the sample's raster) that hold pixel (y, x): interior + halo copies
struct PixSlots {
  int n;
  int o[4];
};
__device__ __forceinline__ PixSlots pixel_slots(const TcGeom& g, int y, int x) {
  int row1 = (y == 0) ? g.H + 1 : ((y == g.H - 1) ? 0 : -1);
  int col0, col1;
  if (g.mode) {
    const int seg = x >> 3, sl = (x & 7) + 1;
    col0 = seg * 10 + sl;
    col1 = (sl == 1) ? ((seg + g.nseg - 1) % g.nseg) * 10 + 9 : ((sl == 8) ? ((seg + 1) % g.nseg) * 10 : -1);
  } else {
    col0 = x + 1;
    col1 = (x == 0) ? g.W + 1 : ((x == g.W - 1) ? 0 : -1);
  }
  PixSlots p;
  p.n = 1;
  p.o[0] = (y + 1) * g.RP + col0;
  p.o[1] = p.o[2] = p.o[3] = 0;
  if (col1 >= 0) p.o[p.n++] = (y + 1) * g.RP + col1;
  if (row1 >= 0) {
    p.o[p.n++] = row1 * g.RP + col0;
    if (col1 >= 0) p.o[p.n++] = row1 * g.RP + col1;
  }
  return p;
}

// store 8 channels (one plane) of one pixel, hi and lo parts, into all of its slots
__device__ __forceinline__ void store_plane(__half* act, const TcGeom& g, int plane, int64_t sample_slot0,
                                            const PixSlots& ps, const float (&t)[8]) {
  uint4 vh, vl;
  split2(t[0], t[1], vh.x, vl.x);
  split2(t[2], t[3], vh.y, vl.y);
  split2(t[4], t[5], vh.z, vl.z);
  split2(t[6], t[7], vh.w, vl.w);
  uint4* base_hi = reinterpret_cast<uint4*>(act) + ((int64_t)((plane >> 1) * 4 + (plane & 1)) * g.slots + sample_slot0);
  uint4* base_lo = base_hi + 2 * g.slots;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (a < ps.n) {
      base_hi[ps.o[a]] = vh;
      base_lo[ps.o[a]] = vl;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight preparation: w [cout][cin][3][3] float32 -> blob [kstep][tap][hi|lo][p2][Np][8] binary16
// ---------------------------------------------------------------------------------------------
struct TcPrepParams {
  const float* params;
  __half* wblob;
  int C, Np, KS, nconv;
  int pair;                      // blob layout of the CTA-pair kernel: [kstep][dy][half][dx][hi|lo][p2][Np/2][8]
  int64_t w_off[kTcMaxLayers];   // offset of the conv weight in `params`
  int64_t b_off[kTcMaxLayers];   // offset of the conv bias in `params`, -1 = no bias
  int64_t blob_off[kTcMaxLayers];
  float* bias_pad;               // [nconv][Np]
  int transpose;                 // backward-data blobs: W'[c][o][tap] = w[o][c][8 - tap], no bias
  int pair_k;                    // last K step: tap-pair slots (see the MMA issuer of resconv_tc2_kernel)
  float* wnorm;                  // transpose: [nconv] max_c sum_{o,tap} |w[o][c][tap]|
};

__global__ void __launch_bounds__(256) tc_weight_prep_kernel(TcPrepParams p) {
  const int L = blockIdx.y;
  const float* w = p.params + p.w_off[L];
  __half* out = p.wblob + p.blob_off[L];
  const int Kp = p.KS * 16;
  if (blockIdx.x == 0)
    for (int o = threadIdx.x; o < p.Np; o += blockDim.x)
      p.bias_pad[L * p.Np + o] = (o < p.C && p.b_off[L] >= 0 && !p.transpose) ? p.params[p.b_off[L] + o] : 0.f;
  const int n = 9 * Kp * p.Np;  // one entry per (tap, c, o)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int o = e % p.Np, c = (e / p.Np) % Kp, tap = e / (p.Np * Kp);
    // source (channel, tap) of blob entry (slot `tap`, K position c).  Last K step of a tap-pair layout: slot
    // (dy', dx') of the stage holds the channel group of taps 2 pi and 2 pi + 1, pi = 2 dy' + dx' (dx' < 2, pi <= 4),
    // in its two K halves
    int cs = c, ts = tap;
    bool live = true;
    if (p.pair_k && (c >> 4) == p.KS - 1) {
      const int dyp = tap / 3, dxp = tap - dyp * 3, pi = dyp * 2 + dxp;
      ts = 2 * pi + ((c >> 3) & 1);
      cs = (p.KS - 1) * 16 + (c & 7);
      live = dxp < 2 && pi <= 4 && ts <= 8;
    }
    float v = 0.f;
    if (live && o < p.C && cs < p.C)
      v = kWScale * (p.transpose ? w[((int64_t)cs * p.C + o) * 9 + (8 - ts)] : w[((int64_t)o * p.C + cs) * 9 + ts]);
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const int ks = c >> 4, p2 = (c >> 3) & 1, j = c & 7;
    if (p.pair) {
      const int nh = p.Np >> 1, half = o / nh, ol = o - half * nh, dy = tap / 3, dx = tap - dy * 3;
      const int64_t base = ((((((int64_t)ks * 3 + dy) * 2 + half) * 3 + dx) * 2 + 0) * 2 + p2) * nh + ol;
      out[base * 8 + j] = h;
      out[(base + 2 * nh) * 8 + j] = l;
    } else {
      const int64_t base = ((((int64_t)ks * 9 + tap) * 2 + 0) * 2 + p2) * p.Np + o;
      out[base * 8 + j] = h;
      out[(base + 2 * p.Np) * 8 + j] = l;
    }
  }
}

// operator-norm bound of the transposed convolutions: block per layer
__global__ void __launch_bounds__(256) tc_wnorm_kernel(TcPrepParams p) {
  __shared__ float red[8];
  const int L = blockIdx.x;
  const float* w = p.params + p.w_off[L];
  float m = 0.f;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float a = 0.f;
    for (int o = 0; o < p.C; ++o)
      for (int tap = 0; tap < 9; ++tap) a += fabsf(w[((int64_t)o * p.C + c) * 9 + tap]);
    m = fmaxf(m, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    p.wnorm[L] = m * 1.001f;
  }
}

// ---------------------------------------------------------------------------------------------
// first layer (conv1 of block 0, cin = 1; conv_nets.py:84-86): CUDA cores, writes the operand raster of
// conv2_0 = split(4 * gelu(conv1_0(s / sqrt 2) + b)) and optionally the raw pre-activation.
// thread per (sample, pixel)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_first_layer_kernel(const int8_t* __restrict__ spins, const float* __restrict__ w,
                                                             const float* __restrict__ b, TcGeom g, int C, int Np,
                                                             __half* __restrict__ act, float* __restrict__ raw_out,
                                                             const long long* __restrict__ ns_dev, int precise) {
  // thread per (sample, pixel): the nine neighbour spins are read once and reused for all channels; the weights
  // [C][9] and biases sit in shared memory (broadcast reads)
  extern __shared__ float wb_s[];  // [Np * 9] weights (zero padded), [Np] biases
  const int N = g.H * g.W, planes = Np >> 3;
  for (int e = threadIdx.x; e < Np * 9; e += blockDim.x) wb_s[e] = (e / 9 < C) ? w[e] : 0.f;
  for (int e = threadIdx.x; e < Np; e += blockDim.x) wb_s[Np * 9 + e] = (e < C) ? b[e] : 0.f;
  __syncthreads();
  const float* bs = wb_s + Np * 9;
  const GeluConst gk = gelu_const(1.0f, precise);
  const int64_t total = (ns_dev ? (int64_t)*ns_dev : g.ns) * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % N);
    const int64_t s = e / N;
    const int y = pix / g.W, x = pix % g.W;
    float sv[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        int yy = y + dy - 1, xx = x + dx - 1;
        yy += (yy < 0) ? g.H : 0; yy -= (yy >= g.H) ? g.H : 0;
        xx += (xx < 0) ? g.W : 0; xx -= (xx >= g.W) ? g.W : 0;
        sv[dy * 3 + dx] = 0.70710678118654752f * (float)spins[s * N + yy * g.W + xx];
      }
    const PixSlots ps = pixel_slots(g, y, x);
    for (int plane = 0; plane < planes; ++plane) {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = plane * 8 + j;
        float v = bs[c];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) v = fmaf(wb_s[c * 9 + tap], sv[tap], v);
        if (raw_out && c < C) raw_out[(s * C + c) * N + pix] = v;
        t[j] = (c < C) ? gelu_scaled(v, gk) : 0.f;
      }
      store_plane(act, g, plane, s * g.Ps, ps, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the persistent tensor-core kernel
// ---------------------------------------------------------------------------------------------
// Epilogue of one plane (8 channels) of one pixel:
//   epi_load : a <- a * 2^-10 + bias + residual
//   epi_store: raw output, gelu, binary16 split, operand stores with halo copies
// PLANAR (forward-only residual stream [ns, Np/8, N, 8]): every buffer is channel-padded, no per-channel guards;
// otherwise raw / res are [C][N] slices (channel stride N) and channels >= nvalid are skipped.
// The layer flags (res / raw null, write_act) are warp-uniform.
template <bool PLANAR>
__device__ __forceinline__ void epi_load(float (&a)[8], float resv, int nvalid, const float* __restrict__ bias,
                                         const float* res, int N, float out_scale) {
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + 1);
  float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (res) {
    if (PLANAR) {
      const float4 r0 = *reinterpret_cast<const float4*>(res);
      const float4 r1 = *reinterpret_cast<const float4*>(res + 4);
      r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) r[j] = res[(int64_t)j * N];
    }
  }
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], out_scale, resv + bb[j]) + r[j];
}

template <bool PLANAR>
__device__ __forceinline__ void epi_store(const float (&v)[8], int nvalid, float* raw, int N, bool write_act,
                                          const GeluConst& gk, uint4* act_hi, int64_t lo_off, const PixSlots& ps) {
  if (raw) {
    if (PLANAR) {
      *reinterpret_cast<float4*>(raw) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(raw + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) raw[(int64_t)j * N] = v[j];
    }
  }
  if (write_act) {
    // channels >= C carry gelu(resv) at most: finite, and multiplied by zero weights in the next layer
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = gelu_scaled(v[j], gk);
    uint4 vh, vl;
    split2(t[0], t[1], vh.x, vl.x);
    split2(t[2], t[3], vh.y, vl.y);
    split2(t[4], t[5], vh.z, vl.z);
    split2(t[6], t[7], vh.w, vl.w);
    uint4* act_lo = act_hi + lo_off;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < ps.n) {
        act_hi[ps.o[q]] = vh;
        act_lo[ps.o[q]] = vl;
      }
  }
}

// PL = planes (8-channel groups) per epilogue thread: Np <= 24 * PL
template <int PL>
__global__ void __launch_bounds__(kTcThreads, 1) resconv_tc_kernel(const __grid_constant__ TcNetParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const TcGeom& g = p.g;
  const uint32_t run_bytes = (uint32_t)g.TS * 16u;           // one (hi|lo, p2) plane of the activation tile
  const uint32_t act_stage_bytes = 4u * run_bytes;
  const uint32_t w_tap_bytes = 4u * (uint32_t)p.Np * 16u;    // [hi|lo][p2][Np][16 B]
  const uint32_t w_stage_bytes = 3u * w_tap_bytes;           // one kernel row (3 taps)
  unsigned char* act_s = smem;
  unsigned char* w_s = act_s + (size_t)p.act_stages * act_stage_bytes;
  uint64_t* act_full = reinterpret_cast<uint64_t*>(w_s + (size_t)p.w_stages * w_stage_bytes);
  uint64_t* act_empty = act_full + p.act_stages;
  uint64_t* w_full = act_empty + p.act_stages;
  uint64_t* w_empty = w_full + p.w_stages;
  uint64_t* tmem_full = w_empty + p.w_stages;   // [1]
  uint64_t* tmem_empty = tmem_full + 1;         // [1]
  uint64_t* act_ready = tmem_empty + 1;         // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(act_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.act_stages; ++i) { mbar_init(act_full + i, 1); mbar_init(act_empty + i, 1); }
    for (int i = 0; i < p.w_stages; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kEpiWarps);
    mbar_init(act_ready + 0, kEpiWarps);
    mbar_init(act_ready + 1, kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // accumulator columns: tile t -> main at (2t) * col_stride, cross terms at (2t + 1) * col_stride
  const uint32_t col_stride = (uint32_t)((p.Np + 31) & ~31);

  const int64_t ns_rt = p.ns_dev ? (int64_t)*p.ns_dev : g.ns;  // runtime sample count
  const int64_t nitems = (ns_rt + g.spi - 1) / g.spi;
  const int64_t nrounds = (nitems + 2 * (int64_t)gridDim.x - 1) / (2 * (int64_t)gridDim.x);
  const int nl = p.layer1 - p.layer0;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      uint32_t as = 0, pa = 0, ws = 0, pw = 0;
      uint32_t cnt[2] = {0, 0};
      unsigned long long t_ready = 0, t_aempty = 0, t_wempty = 0;
      const long long t_begin = clock64();
      for (int64_t r = 0; r < nrounds; ++r)
        for (int li = 0; li < nl; ++li)
          for (int j = 0; j < 2; ++j) {
            const int64_t item = (2 * r + j) * gridDim.x + blockIdx.x;
            if (item >= nitems) continue;
            if (cnt[j] > 0) QTX_TIMED_WAIT(t_ready, act_ready + j, (cnt[j] - 1) & 1);  // the previous layer of this item is stored
            ++cnt[j];
            const int64_t wblob_off = p.layer[p.layer0 + li].wblob_off;
            const int64_t slot0 = item * g.spi * g.Ps;
            for (int ks = 0; ks < p.KS; ++ks) {
              QTX_TIMED_WAIT(t_aempty, act_empty + as, pa ^ 1);
              mbar_expect_tx(act_full + as, act_stage_bytes);
              unsigned char* dst = act_s + (size_t)as * act_stage_bytes;
#pragma unroll
              for (int run = 0; run < 4; ++run)
                bulk_g2s(dst + (size_t)run * run_bytes, p.act + (((int64_t)ks * 4 + run) * g.slots + slot0) * 8, run_bytes,
                         act_full + as);
              if (++as == (uint32_t)p.act_stages) { as = 0; pa ^= 1; }
              for (int dy = 0; dy < 3; ++dy) {
                QTX_TIMED_WAIT(t_wempty, w_empty + ws, pw ^ 1);
                mbar_expect_tx(w_full + ws, w_stage_bytes);
                bulk_g2s(w_s + (size_t)ws * w_stage_bytes,
                         p.wblob + wblob_off + ((int64_t)ks * 9 + dy * 3) * (w_tap_bytes / 2), w_stage_bytes, w_full + ws);
                if (++ws == (uint32_t)p.w_stages) { ws = 0; pw ^= 1; }
              }
            }
          }
      if (p.dbg) {
        unsigned long long* d = p.dbg + (size_t)blockIdx.x * 16;
        d[0] = (unsigned long long)(clock64() - t_begin); d[1] = t_ready; d[2] = t_aempty; d[3] = t_wempty;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =====
    {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // D=F32, A=B=F16, K-major
      // descriptors (make_desc_nosw) split into words; everything below is in 16-byte units
      const uint32_t a_hi_w = (uint32_t)g.SB | (1u << 14);          // SBO = SB slots, version 1
      const uint32_t b_hi_w = (128u >> 4) | (1u << 14);             // SBO = 8 rows x 16 B
      const uint32_t run16 = run_bytes >> 4, np16 = (uint32_t)p.Np; // LBO of A (plane pitch) and B
      const uint32_t act_base = (smem_u32(act_s) >> 4) | (run16 << 16);
      const uint32_t w_base = (smem_u32(w_s) >> 4) | (np16 << 16);
      const uint32_t act_stage16 = act_stage_bytes >> 4, w_stage16 = w_stage_bytes >> 4, w_tap16 = w_tap_bytes >> 4;
      // per-tile slot offset inside the staged activation tile
      const uint32_t tile_off0 = 0u;
      const uint32_t tile_off1 = (g.tps == 1) ? (uint32_t)g.Ps : (uint32_t)(16 * g.SB);
      const uint32_t d_main0 = tmem_base, d_cross0 = tmem_base + col_stride;
      const uint32_t d_main1 = tmem_base + 2 * col_stride, d_cross1 = tmem_base + 3 * col_stride;
      uint32_t as = 0, pa = 0, ws = 0, pw = 0, q = 0;
      unsigned long long t_tempty = 0, t_afull = 0, t_wfull = 0;
      const long long t_begin = clock64();
      for (int64_t r = 0; r < nrounds; ++r)
        for (int li = 0; li < nl; ++li)
          for (int j = 0; j < 2; ++j) {
            const int64_t item = (2 * r + j) * gridDim.x + blockIdx.x;
            if (item >= nitems) continue;
            QTX_TIMED_WAIT(t_tempty, tmem_empty, (q & 1) ^ 1);  // the epilogue has drained the previous item's accumulators
            ++q;
            tc_fence_after();
            for (int ks = 0; ks < p.KS; ++ks) {
              QTX_TIMED_WAIT(t_afull, act_full + as, pa);
              const uint32_t a_stage = act_base + as * act_stage16;
              for (int dy = 0; dy < 3; ++dy) {
                QTX_TIMED_WAIT(t_wfull, w_full + ws, pw);
                tc_fence_after();
                const uint32_t w_stage = w_base + ws * w_stage16;
                const uint32_t a_row0 = a_stage + tile_off0 + (uint32_t)(dy * g.RP);
                const uint32_t a_row1 = a_stage + tile_off1 + (uint32_t)(dy * g.RP);
                const uint32_t acc0 = (ks == 0 && dy == 0) ? 0u : 1u;
                if (elect_one()) {
#pragma unroll
                  for (int dx = 0; dx < 3; ++dx) {
                    const uint32_t b_h = w_stage + (uint32_t)dx * w_tap16, b_l = b_h + 2u * np16;
                    const uint32_t acc = (dx == 0) ? acc0 : 1u;
                    const uint32_t a0h = a_row0 + dx, a0l = a0h + 2u * run16;
                    const uint32_t a1h = a_row1 + dx, a1l = a1h + 2u * run16;
                    umma_f16_split(d_main0, a0h, a_hi_w, b_h, b_hi_w, idesc, acc);
                    umma_f16_split(d_main1, a1h, a_hi_w, b_h, b_hi_w, idesc, acc);
                    umma_f16_split(d_cross0, a0l, a_hi_w, b_h, b_hi_w, idesc, acc);
                    umma_f16_split(d_cross1, a1l, a_hi_w, b_h, b_hi_w, idesc, acc);
                    umma_f16_split(d_cross0, a0h, a_hi_w, b_l, b_hi_w, idesc, 1u);
                    umma_f16_split(d_cross1, a1h, a_hi_w, b_l, b_hi_w, idesc, 1u);
                  }
                  umma_commit(w_empty + ws);
                  if (dy == 2) umma_commit(act_empty + as);
                  if (dy == 2 && ks == p.KS - 1) umma_commit(tmem_full);
                }
                __syncwarp();
                if (++ws == (uint32_t)p.w_stages) { ws = 0; pw ^= 1; }
              }
              if (++as == (uint32_t)p.act_stages) { as = 0; pa ^= 1; }
            }
          }
      if (p.dbg && lane == 0) {
        unsigned long long* d = p.dbg + (size_t)blockIdx.x * 16;
        d[4] = (unsigned long long)(clock64() - t_begin); d[5] = t_tempty; d[6] = t_afull; d[7] = t_wfull;
      }
    }
  } else {
    // ===== epilogue: warp -> TMEM lane quarter (warp % 4), column group (warp - 2) / 4 owns planes cgp, cgp + 3, ... =====
    const int lq = warp & 3, cgp = (warp - 2) >> 2;
    const int N = g.H * g.W, planes = p.Np >> 3;
    const int m = lq * 32 + lane;
    // M row of this thread -> pixel and operand slots; the same for every item and layer
    bool rvalid[2];
    int rpix[2];
    PixSlots rps[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int tis = (g.tps == 1) ? 0 : t;
      const int slot = g.RP + 1 + tis * 16 * g.SB + (m >> 3) * g.SB + (m & 7);  // within the sample raster
      const int row = slot / g.RP, col = slot - row * g.RP;
      const int y = row - 1;
      int x;
      bool okx;
      if (g.mode) {
        const int seg = col / 10, c10 = col - seg * 10;
        x = seg * 8 + c10 - 1;
        okx = c10 >= 1 && c10 <= 8;
      } else {
        x = col - 1;
        okx = x >= 0 && x < g.W;
      }
      rvalid[t] = y >= 0 && y < g.H && okx;
      rpix[t] = rvalid[t] ? y * g.W + x : 0;
      rps[t] = pixel_slots(g, rvalid[t] ? y : 0, rvalid[t] ? x : 0);
    }
    uint32_t q = 0;
    unsigned long long t_tfull = 0, t_drain = 0;
    const long long t_begin = clock64();
    for (int64_t r = 0; r < nrounds; ++r)
      for (int li = 0; li < nl; ++li)
        for (int j = 0; j < 2; ++j) {
          const int64_t item = (2 * r + j) * gridDim.x + blockIdx.x;
          if (item >= nitems) continue;
          const TcLayer& L = p.layer[p.layer0 + li];
          if (lane == 0) QTX_TIMED_WAIT(t_tfull, tmem_full, q & 1);
          __syncwarp();
          const long long t_d0 = clock64();
          ++q;
          tc_fence_after();
          // ---- drain: main + cross accumulators of both tiles into registers, then release TMEM ----
          float acc[2][PL][8];
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int k = 0; k < PL; ++k) {
              const int plane = cgp + kColGroups * k;
              if (plane < planes) {
                uint32_t rm[8], rc[8];
                const uint32_t a = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(2 * t) * col_stride + plane * 8;
                tmem_ld8_nowait(a, rm);
                tmem_ld8_nowait(a + col_stride, rc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[t][k][jj] = __uint_as_float(rm[jj]) + __uint_as_float(rc[jj]);
              }
            }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
          t_drain += (unsigned long long)(clock64() - t_d0);
          // ---- bias / residual / raw output / gelu / split / operand store ----
          const GeluConst gk = gelu_const(L.out_alpha, p.precise);
          const float* Lbias = L.bias;
          const float* Lres = L.res;
          float* Lraw = L.raw_out;
          const int8_t* Lspin = L.res_spin;
          const bool Lwrite = L.write_act != 0;
          auto run = [&](auto planar_tag) {
            constexpr bool PLANAR = decltype(planar_tag)::value;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int64_t s = item * g.spi + ((g.tps == 1) ? t : 0);
              if (!rvalid[t] || s >= ns_rt) continue;
              const int pix = rpix[t];
              const float resv = Lspin ? (float)Lspin[s * N + pix] : 0.f;
              uint4* act_sample = reinterpret_cast<uint4*>(p.act) + s * g.Ps;
              const int64_t raw_sample = PLANAR ? s * planes * N * 8 + (int64_t)pix * 8 : s * p.C * N + pix;
#pragma unroll
              for (int k = 0; k < PL; ++k) {
                const int plane = cgp + kColGroups * k;
                if (plane >= planes) continue;
                const int c0 = plane * 8;
                const int64_t roff = raw_sample + (PLANAR ? (int64_t)plane * N * 8 : (int64_t)c0 * N);
                uint4* act_hi = act_sample + (int64_t)((plane >> 1) * 4 + (plane & 1)) * g.slots;
                if (PLANAR && Lres && plane + kColGroups < planes)  // next plane's residual -> L1 while this one computes
                  prefetch_l1(Lres + roff + (int64_t)kColGroups * N * 8);
                epi_load<PLANAR>(acc[t][k], resv, p.C - c0, Lbias + c0, Lres ? Lres + roff : nullptr, N, p.out_scale);
                epi_store<PLANAR>(acc[t][k], p.C - c0, Lraw ? Lraw + roff : nullptr, N, Lwrite, gk, act_hi, 2 * g.slots,
                                  rps[t]);
              }
            }
          };
          if (L.planar) run(std::true_type{});
          else run(std::false_type{});
          fence_proxy_async_global();
          __syncwarp();
          if (lane == 0) mbar_arrive(act_ready + j);
        }
    if (p.dbg && warp == 2 && lane == 0) {
      unsigned long long* d = p.dbg + (size_t)blockIdx.x * 16;
      d[8] = (unsigned long long)(clock64() - t_begin); d[9] = t_tfull; d[10] = t_drain;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cluster 2x1x1, tcgen05.mma.cta_group::2, UMMA M = 256).  A work item is FOUR 128-row
// tiles: pair-tile t in {0, 1} = rows of CTA 0 (tile 2t) and CTA 1 (tile 2t + 1).  Each CTA stages only the
// activation slots of its own tiles and HALF of the weight rows (N/2 out-channels), so the shared-memory
// operand reads per MMA drop from 4 + 3 KB to 4 + 1.5 KB per CTA (the single-CTA kernel is bound by exactly
// these reads) and the weight stream from L2 is halved.
//   - both CTAs run a TMA producer (cp.async.bulk.tensor ... .cta_group::2) signalling the LEADER's full barriers
//   - the leader's elected lane issues the MMAs; tcgen05.commit ... multicast::cluster frees the stages and
//     publishes the accumulators in both CTAs
//   - epilogue warps of both CTAs drain their own TMEM lanes; operand stores of a sample cross the CTA boundary
//     (halo rows), so "previous layer stored" is signalled to BOTH producers (remote mbarrier arrive, cluster scope)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc2_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tc2_mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tc2_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// no memory ordering: used where the data hand-over is TMEM (ordered by tcgen05.wait::ld + fence::before_thread_sync)
__device__ __forceinline__ void tc2_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool tc2_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tc2_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!tc2_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 27)) __trap();
  }
}
// TMA loads of a CTA pair: complete_tx goes to the barrier at the same offset in the LEADER CTA
__device__ __forceinline__ void tc2_tma_3d(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                           int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc2_tma_2d(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc2_tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc2_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc2_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc2_commit(uint64_t* bar) {  // arrives on `bar` of BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

// BWD: the tower of backward-data layers (mode 1) -- a separate instantiation so that its epilogue does not cost the
// forward kernel registers
template <int PL, bool BWD>
__global__ void __launch_bounds__(kTcThreads, 1)
    resconv_tc2_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapW,
                       const __grid_constant__ TcNetParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const TcGeom& g = p.g;
  const int Nh = p.Np >> 1;                                   // out-channels staged by this CTA
  const uint32_t run_bytes = (uint32_t)g.TSh * 16u;           // one (hi|lo, p2) plane of one tile's slots
  const uint32_t tile_bytes = 4u * run_bytes;
  const uint32_t act_stage_bytes = 2u * tile_bytes;           // both pair-tiles
  const uint32_t w_tap_bytes = 4u * (uint32_t)Nh * 16u;       // [hi|lo][p2][Nh][16 B]
  const uint32_t w_stage_bytes = 3u * w_tap_bytes;
  unsigned char* act_s = smem;
  unsigned char* w_s = act_s + (size_t)p.act_stages * act_stage_bytes;
  uint64_t* act_full = reinterpret_cast<uint64_t*>(w_s + (size_t)p.w_stages * w_stage_bytes);
  uint64_t* act_empty = act_full + p.act_stages;
  uint64_t* w_full = act_empty + p.act_stages;
  uint64_t* w_empty = w_full + p.w_stages;
  uint64_t* tmem_full = w_empty + p.w_stages;   // [1]
  uint64_t* tmem_empty = tmem_full + 1;         // [1]  (leader's is used)
  uint64_t* act_ready = tmem_empty + 1;         // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(act_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.act_stages; ++i) { mbar_init(act_full + i, 1); mbar_init(act_empty + i, 1); }
    for (int i = 0; i < p.w_stages; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 2 * kEpiWarps);
    mbar_init(act_ready + 0, 2);  // one signalling thread per CTA
    mbar_init(act_ready + 1, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tc2_tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc2_cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t col_stride = (uint32_t)((p.Np + 31) & ~31);

  const int64_t ns_rt = p.ns_dev ? (int64_t)*p.ns_dev : g.ns;  // runtime sample count
  const int64_t nitems = (ns_rt + g.spi - 1) / g.spi;
  const int npairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const int64_t nrounds = (nitems + 2 * (int64_t)npairs - 1) / (2 * (int64_t)npairs);
  const int nl = p.layer1 - p.layer0;
  // this CTA's tile of pair-tile t: tile index 2t + rank -> (sample in item, tile in sample)
  int tsl[2], ttis[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int q = 2 * t + (int)rank;
    tsl[t] = q / g.tps;
    ttis[t] = q - tsl[t] * g.tps;
  }

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      uint32_t as = 0, pa = 0, ws = 0, pw = 0;
      uint32_t cnt[2] = {0, 0};
      for (int64_t r = 0; r < nrounds; ++r)
        for (int li = 0; li < nl; ++li)
          for (int j = 0; j < 2; ++j) {
            const int64_t item = (2 * r + j) * npairs + pair_id;
            if (item >= nitems) continue;
            if (cnt[j] > 0) {
              tc2_wait_cluster(act_ready + j, (cnt[j] - 1) & 1);  // both CTAs stored the previous layer of this item
              fence_proxy_async_global();
            }
            ++cnt[j];
            const int layer = p.layer0 + li;
            const int row_in = p.layer[layer].in_buf * p.KS * 4;
            const int64_t slot0 = item * g.spi * g.Ps;
            for (int ks = 0; ks < p.KS; ++ks) {
              mbar_wait(act_empty + as, pa ^ 1);
              const uint32_t lbar = smem_u32(act_full + as) & 0xFEFFFFFFu;
              // tap-pair K step: only the plane with the real channels (p2 = 0) is read by the MMAs
              const bool half_step = p.pair_k != 0 && p.keep_pad == 0 && ks == p.KS - 1;
              if (leader) mbar_expect_tx(act_full + as, half_step ? act_stage_bytes : 2u * act_stage_bytes);
              unsigned char* dst = act_s + (size_t)as * act_stage_bytes;
              // the map counts 8-byte elements (2 per slot); one box = half of one plane run (<= 256 elements)
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const int e0 = 2 * (int)(slot0 + (int64_t)tsl[t] * g.Ps + ttis[t] * 16 * g.SB);
#pragma unroll
                for (int run = 0; run < 4; ++run) {
                  if (half_step && (run & 1)) continue;
                  unsigned char* d = dst + (size_t)t * tile_bytes + (size_t)run * run_bytes;
                  tc2_tma_2d(d, &tmapA, lbar, e0, row_in + ks * 4 + run);
                  tc2_tma_2d(d + (run_bytes >> 1), &tmapA, lbar, e0 + g.TSh, row_in + ks * 4 + run);
                }
              }
              if (++as == (uint32_t)p.act_stages) { as = 0; pa ^= 1; }
              for (int dy = 0; dy < 3; ++dy) {
                mbar_wait(w_empty + ws, pw ^ 1);
                const uint32_t wbar = smem_u32(w_full + ws) & 0xFEFFFFFFu;
                if (leader) mbar_expect_tx(w_full + ws, 2u * w_stage_bytes);
                // blob [layer][kstep][dy][half] -> one contiguous stage = 6 map rows of 4 * Nh 8-byte elements
                const int row0 = (((layer * p.KS + ks) * 3 + dy) * 2 + (int)rank) * 6;
                tc2_tma_2d(w_s + (size_t)ws * w_stage_bytes, &tmapW, wbar, 0, row0);
                if (++ws == (uint32_t)p.w_stages) { ws = 0; pw ^= 1; }
              }
            }
          }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only; the whole warp runs the loop, one elected lane issues =====
    if (leader) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);  // M = 256
      const uint32_t a_hi_w = (uint32_t)g.SB | (1u << 14);
      const uint32_t b_hi_w = (128u >> 4) | (1u << 14);
      const uint32_t run16 = run_bytes >> 4, nh16 = (uint32_t)Nh;
      const uint32_t act_base = (smem_u32(act_s) >> 4) | (run16 << 16);
      const uint32_t w_base = (smem_u32(w_s) >> 4) | (nh16 << 16);
      const uint32_t act_stage16 = act_stage_bytes >> 4, w_stage16 = w_stage_bytes >> 4, w_tap16 = w_tap_bytes >> 4;
      const uint32_t tile16 = tile_bytes >> 4;
      const uint32_t d_main0 = tmem_base, d_cross0 = tmem_base + col_stride;
      const uint32_t d_main1 = tmem_base + 2 * col_stride, d_cross1 = tmem_base + 3 * col_stride;
      uint32_t as = 0, pa = 0, ws = 0, pw = 0, q = 0;
      unsigned long long t_tempty = 0, t_afull = 0, t_wfull = 0;
      const long long t_begin = clock64();
      for (int64_t r = 0; r < nrounds; ++r)
        for (int li = 0; li < nl; ++li)
          for (int j = 0; j < 2; ++j) {
            const int64_t item = (2 * r + j) * npairs + pair_id;
            if (item >= nitems) continue;
            {
              const long long _t0 = clock64();
              tc2_wait_cluster(tmem_empty, (q & 1) ^ 1);  // both CTAs' epilogues drained the previous item
              t_tempty += (unsigned long long)(clock64() - _t0);
            }
            ++q;
            tc_fence_after();
            for (int ks = 0; ks < p.KS; ++ks) {
              QTX_TIMED_WAIT(t_afull, act_full + as, pa);
              const uint32_t a_stage = act_base + as * act_stage16;
              for (int dy = 0; dy < 3; ++dy) {
                QTX_TIMED_WAIT(t_wfull, w_full + ws, pw);
                tc_fence_after();
                const uint32_t w_stage = w_base + ws * w_stage16;
                const uint32_t a_row0 = a_stage + (uint32_t)(dy * g.RP);
                const uint32_t a_row1 = a_row0 + tile16;
                const uint32_t acc0 = (ks == 0 && dy == 0) ? 0u : 1u;
                const bool pair_step = p.pair_k != 0 && ks == p.KS - 1;  // warp-uniform
                if (elect_one()) {
                  if (pair_step) {
                    // The last K step holds at most 8 channels: instead of nine MMAs with a half-empty K, one MMA
                    // takes that channel group of TWO taps -- its second K core matrix is the partner tap's, reached
                    // through the leading-dimension offset of the descriptor.  Stage dy holds the tap pairs
                    // (0,1),(2,3) | (4,5),(6,7) | (8,-) (tc_weight_prep_kernel); 5 MMA groups instead of 9.
                    const uint32_t a_addr = (smem_u32(act_s) >> 4) + as * act_stage16;
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                      const int pi = dy * 2 + sl;
                      if (pi <= 4) {
                        const int t0 = 2 * pi, t1 = t0 + 1;
                        const uint32_t off0 = (uint32_t)((t0 / 3) * g.RP + t0 % 3);
                        // tap 8 has no partner: its second K half (zero weights) re-reads the first, which is finite --
                        // one slot further may lie behind the last sample's raster, where nothing is ever written
                        const uint32_t off1 = t1 <= 8 ? (uint32_t)((t1 / 3) * g.RP + t1 % 3) : off0;
                        const uint32_t b_h = w_stage + (uint32_t)sl * w_tap16, b_l = b_h + 2u * nh16;
                        const uint32_t acc = (sl == 0) ? acc0 : 1u;
                        const uint32_t a0h = (a_addr + off0) | ((off1 - off0) << 16), a0l = a0h + 2u * run16;
                        const uint32_t a1h = a0h + tile16, a1l = a1h + 2u * run16;
                        tc2_umma(d_main0, a0h, a_hi_w, b_h, b_hi_w, idesc, acc);
                        tc2_umma(d_main1, a1h, a_hi_w, b_h, b_hi_w, idesc, acc);
                        tc2_umma(d_cross0, a0l, a_hi_w, b_h, b_hi_w, idesc, acc);
                        tc2_umma(d_cross1, a1l, a_hi_w, b_h, b_hi_w, idesc, acc);
                        tc2_umma(d_cross0, a0h, a_hi_w, b_l, b_hi_w, idesc, 1u);
                        tc2_umma(d_cross1, a1h, a_hi_w, b_l, b_hi_w, idesc, 1u);
                      }
                    }
                  } else {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                      const uint32_t b_h = w_stage + (uint32_t)dx * w_tap16, b_l = b_h + 2u * nh16;
                      const uint32_t acc = (dx == 0) ? acc0 : 1u;
                      const uint32_t a0h = a_row0 + dx, a0l = a0h + 2u * run16;
                      const uint32_t a1h = a_row1 + dx, a1l = a1h + 2u * run16;
                      tc2_umma(d_main0, a0h, a_hi_w, b_h, b_hi_w, idesc, acc);
                      tc2_umma(d_main1, a1h, a_hi_w, b_h, b_hi_w, idesc, acc);
                      tc2_umma(d_cross0, a0l, a_hi_w, b_h, b_hi_w, idesc, acc);
                      tc2_umma(d_cross1, a1l, a_hi_w, b_h, b_hi_w, idesc, acc);
                      tc2_umma(d_cross0, a0h, a_hi_w, b_l, b_hi_w, idesc, 1u);
                      tc2_umma(d_cross1, a1h, a_hi_w, b_l, b_hi_w, idesc, 1u);
                    }
                  }
                  tc2_commit(w_empty + ws);
                  if (dy == 2) tc2_commit(act_empty + as);
                  if (dy == 2 && ks == p.KS - 1) tc2_commit(tmem_full);
                }
                __syncwarp();
                if (++ws == (uint32_t)p.w_stages) { ws = 0; pw ^= 1; }
              }
              if (++as == (uint32_t)p.act_stages) { as = 0; pa ^= 1; }
            }
          }
      if (p.dbg && lane == 0) {
        unsigned long long* d = p.dbg + (size_t)pair_id * 16;
        d[4] = (unsigned long long)(clock64() - t_begin); d[5] = t_tempty; d[6] = t_afull; d[7] = t_wfull;
      }
    }
  } else {
    // ===== epilogue (both CTAs): warp -> TMEM lane quarter (warp % 4), column group (warp - 2) / 4 =====
    const int lq = warp & 3, cgp = (warp - 2) >> 2;
    const int N = g.H * g.W, planes = p.Np >> 3;
    const int m = lq * 32 + lane;
    bool rvalid[2];
    int rpix[2];
    PixSlots rps[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int slot = g.RP + 1 + ttis[t] * 16 * g.SB + (m >> 3) * g.SB + (m & 7);  // within the sample raster
      const int row = slot / g.RP, col = slot - row * g.RP;
      const int y = row - 1;
      int x;
      bool okx;
      if (g.mode) {
        const int seg = col / 10, c10 = col - seg * 10;
        x = seg * 8 + c10 - 1;
        okx = c10 >= 1 && c10 <= 8;
      } else {
        x = col - 1;
        okx = x >= 0 && x < g.W;
      }
      rvalid[t] = y >= 0 && y < g.H && okx;
      rpix[t] = rvalid[t] ? y * g.W + x : 0;
      rps[t] = pixel_slots(g, rvalid[t] ? y : 0, rvalid[t] ? x : 0);
    }
    // first and last valid pixel of this warp's rows (prefetch lines of the backward-data layers)
    int wpix_lo[2], wpix_hi[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      int lo = rvalid[t] ? rpix[t] : 0x7fffffff, hi = rvalid[t] ? rpix[t] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(FULL, lo, o));
        hi = max(hi, __shfl_xor_sync(FULL, hi, o));
      }
      wpix_lo[t] = lo == 0x7fffffff ? 0 : lo;
      wpix_hi[t] = hi;
    }
    const uint32_t tmem_empty_leader = tc2_mapa(smem_u32(tmem_empty), 0);
    uint32_t ready_addr[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) ready_addr[j][c] = tc2_mapa(smem_u32(act_ready + j), (uint32_t)c);
    uint32_t q = 0;
    unsigned long long t_tfull = 0, t_drain = 0;
    const long long t_begin = clock64();
    for (int64_t r = 0; r < nrounds; ++r)
      for (int li = 0; li < nl; ++li)
        for (int j = 0; j < 2; ++j) {
          const int64_t item = (2 * r + j) * npairs + pair_id;
          if (item >= nitems) continue;
          const TcLayer& L = p.layer[p.layer0 + li];
          if (L.planar && L.res) {  // first residual plane of both tiles -> L1 while the MMAs finish
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int64_t s = item * g.spi + tsl[t];
              if (rvalid[t] && s < ns_rt && cgp < planes)
                prefetch_l1(L.res + s * planes * N * 8 + ((int64_t)cgp * N + rpix[t]) * 8);
            }
          }
          if constexpr (BWD) {
            // backward-data layer: the raw pre-activations and the residual gradient come from HBM (written by earlier
            // launches / layers); pull this warp's lines into L2 while the MMAs of the item run.  Lane -> channel.
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int64_t s = item * g.spi + tsl[t];
              if (s >= ns_rt) continue;
              for (int idx = lane; idx < PL * 8; idx += 32) {
                const int c = (cgp + kColGroups * (idx >> 3)) * 8 + (idx & 7);
                if (c < p.C) {
                  const int64_t o0 = (s * p.C + c) * N;
                  prefetch_l2(L.mul + o0 + wpix_lo[t]);
                  prefetch_l2(L.mul + o0 + wpix_hi[t]);
                  if (L.res) {
                    prefetch_l2(L.res + o0 + wpix_lo[t]);
                    prefetch_l2(L.res + o0 + wpix_hi[t]);
                  }
                }
              }
            }
          }
          if (lane == 0) QTX_TIMED_WAIT(t_tfull, tmem_full, q & 1);
          __syncwarp();
          const long long t_d0 = clock64();
          ++q;
          tc_fence_after();
          float acc[2][PL][8];
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int k = 0; k < PL; ++k) {
              const int plane = cgp + kColGroups * k;
              if (plane < planes) {
                uint32_t rm[8], rc[8];
                const uint32_t a = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(2 * t) * col_stride + plane * 8;
                tmem_ld8_nowait(a, rm);
                tmem_ld8_nowait(a + col_stride, rc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[t][k][jj] = __uint_as_float(rm[jj]) + __uint_as_float(rc[jj]);
              }
            }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) tc2_arrive_remote_relaxed(tmem_empty_leader);
          t_drain += (unsigned long long)(clock64() - t_d0);
          const GeluConst gk = gelu_const(L.out_alpha, p.precise);
          const float* Lbias = L.bias;
          const float* Lres = L.res;
          float* Lraw = L.raw_out;
          const int8_t* Lspin = L.res_spin;
          const bool Lwrite = L.write_act != 0;
          auto run = [&](auto planar_tag) {
            constexpr bool PLANAR = decltype(planar_tag)::value;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int64_t s = item * g.spi + tsl[t];
              if (!rvalid[t] || s >= ns_rt) continue;
              const int pix = rpix[t];
              const float resv = Lspin ? (float)Lspin[s * N + pix] : 0.f;
              uint4* act_sample = reinterpret_cast<uint4*>(p.act) + (int64_t)L.out_buf * p.buf_u4 + s * g.Ps;
              const int64_t raw_sample = PLANAR ? s * planes * N * 8 + (int64_t)pix * 8 : s * p.C * N + pix;
#pragma unroll
              for (int k = 0; k < PL; ++k) {
                const int plane = cgp + kColGroups * k;
                if (plane >= planes) continue;
                const int c0 = plane * 8;
                const int64_t roff = raw_sample + (PLANAR ? (int64_t)plane * N * 8 : (int64_t)c0 * N);
                uint4* act_hi = act_sample + (int64_t)((plane >> 1) * 4 + (plane & 1)) * g.slots;
                if (PLANAR && Lres && plane + kColGroups < planes)  // next plane's residual -> L1 while this one computes
                  prefetch_l1(Lres + roff + (int64_t)kColGroups * N * 8);
                epi_load<PLANAR>(acc[t][k], resv, p.C - c0, Lbias + c0, Lres ? Lres + roff : nullptr, N, p.out_scale);
                epi_store<PLANAR>(acc[t][k], p.C - c0, Lraw ? Lraw + roff : nullptr, N, Lwrite, gk, act_hi, 2 * g.slots,
                                  rps[t]);
              }
            }
          };
          // The two layer kinds of the forward-only tower (sweep, Oloc) with the pointers formed once per tile: IO =
          // planar residual stream in / out (conv2 layers), !IO = operand store only (conv1 layers).  Same arithmetic
          // as `run`; the generic version recomputed the 64-bit element offsets of the residual, the raw output and
          // every operand store from indices (14 integer instructions per element, at half the FP32 rate).
          // with the tap-pair K step nothing reads the padded channel group behind C (operand plane and residual-stream
          // plane ceil(C/8) .. Np/8 - 1): the forward-only bodies skip it
          const int planes_live = (p.pair_k && !p.keep_pad) ? ((p.C + 7) >> 3) : planes;
          auto run_fast = [&](auto io_tag) {
            constexpr bool IO = decltype(io_tag)::value;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int64_t s = item * g.spi + tsl[t];
              if (!rvalid[t] || s >= ns_rt) continue;
              const int pix = rpix[t];
              float resv = 0.f;
              const float* res_p = nullptr;
              float* raw_p = nullptr;
              if constexpr (IO) {
                resv = Lspin ? (float)Lspin[s * N + pix] : 0.f;
                const int64_t e0 = ((s * planes + cgp) * (int64_t)N + pix) * 8;  // plane cgp of this pixel
                res_p = Lres ? Lres + e0 : nullptr;
                raw_p = Lraw ? Lraw + e0 : nullptr;
              }
              const int64_t rstep = (int64_t)kColGroups * N * 8;  // floats between this thread's planes
              char* act_t = reinterpret_cast<char*>(p.act) + ((int64_t)L.out_buf * p.buf_u4 + s * g.Ps) * 16;
              const int64_t plane_bytes = (int64_t)g.slots * 16, lo_bytes = 2 * plane_bytes;
              int64_t ob[4];
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) ob[q4] = (int64_t)rps[t].o[q4] * 16;
              const int nslot = rps[t].n;
#pragma unroll
              for (int k = 0; k < PL; ++k) {
                const int plane = cgp + kColGroups * k;
                if (plane >= planes_live) continue;
                float(&a)[8] = acc[t][k];
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(Lbias + plane * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(Lbias + plane * 8) + 1);
                float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if constexpr (IO) {
                  if (res_p) {
                    const float4 r0 = *reinterpret_cast<const float4*>(res_p + k * rstep);
                    const float4 r1 = *reinterpret_cast<const float4*>(res_p + k * rstep + 4);
                    r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
                    if (plane + kColGroups < planes) prefetch_l1(res_p + (k + 1) * rstep);
                  }
                }
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) a[jj] = fmaf(a[jj], p.out_scale, resv + bb[jj]) + r[jj];
                if constexpr (IO) {
                  if (raw_p) {
                    *reinterpret_cast<float4*>(raw_p + k * rstep) = make_float4(a[0], a[1], a[2], a[3]);
                    *reinterpret_cast<float4*>(raw_p + k * rstep + 4) = make_float4(a[4], a[5], a[6], a[7]);
                  }
                }
                if (Lwrite) {
                  float tt[8];
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) tt[jj] = gelu_scaled(a[jj], gk);
                  uint4 vh, vl;
                  split2(tt[0], tt[1], vh.x, vl.x);
                  split2(tt[2], tt[3], vh.y, vl.y);
                  split2(tt[4], tt[5], vh.z, vl.z);
                  split2(tt[6], tt[7], vh.w, vl.w);
                  char* ph = act_t + (int64_t)((plane >> 1) * 4 + (plane & 1)) * plane_bytes;
                  *reinterpret_cast<uint4*>(ph + ob[0]) = vh;
                  *reinterpret_cast<uint4*>(ph + lo_bytes + ob[0]) = vl;
                  if (nslot >= 2) {  // halo copies: a pixel has 1, 2 (edge) or 4 (corner) slots
                    *reinterpret_cast<uint4*>(ph + ob[1]) = vh;
                    *reinterpret_cast<uint4*>(ph + lo_bytes + ob[1]) = vl;
                    if (nslot == 4) {
                      *reinterpret_cast<uint4*>(ph + ob[2]) = vh;
                      *reinterpret_cast<uint4*>(ph + lo_bytes + ob[2]) = vl;
                      *reinterpret_cast<uint4*>(ph + ob[3]) = vh;
                      *reinterpret_cast<uint4*>(ph + lo_bytes + ob[3]) = vl;
                    }
                  }
                }
              }
            }
          };
          // backward-data layer: v = conv^T(g) * alpha gelu'(alpha * mul) + res; per-sample scales (see TcLayer)
          const int planes_live_b = (p.pair_k && !p.keep_pad && !p.act_z) ? ((p.C + 7) >> 3) : planes;
          auto run_bwd = [&]() {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int64_t s = item * g.spi + tsl[t];
              if (s >= ns_rt) continue;  // warp-uniform
              const float s_in = __ldcg(L.sig_in + s);
              const float m_in = __ldcg(L.max_in + s);
              const float m_res = L.max_res ? __ldcg(L.max_res + s) : 0.f;
              const float s_out = grad_scale(fmaf(m_in * __ldg(L.wnorm), 1.13f * L.mul_alpha, m_res));
              const float oscale = p.out_scale * kActScale / s_in;  // out_scale = comp / (kActScale kWScale)
              const float ma = L.mul_alpha;
              float vmax = 0.f;
              if (rvalid[t]) {
                const int pix = rpix[t];
                uint4* act_sample = reinterpret_cast<uint4*>(p.act) + (int64_t)L.out_buf * p.buf_u4 + s * g.Ps;
                const int64_t raw_sample = s * p.C * N + pix;
#pragma unroll
                for (int k = 0; k < PL; ++k) {
                  const int plane = cgp + kColGroups * k;
                  if (plane >= planes_live_b) continue;  // nothing reads the padded channel group (tap-pair K step)
                  const int c0 = plane * 8, nvalid = p.C - c0;
                  const int64_t roff = raw_sample + (int64_t)c0 * N;
                  float v[8];
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) {
                    v[jj] = 0.f;
                    if (jj < nvalid) {
                      const float mv = L.mul[roff + (int64_t)jj * N];
                      const float r = Lres ? Lres[roff + (int64_t)jj * N] : 0.f;
                      v[jj] = fmaf(acc[t][k][jj] * oscale, ma * gelu_grad_fast(ma * mv, p.precise), r);
                      if (Lraw) Lraw[roff + (int64_t)jj * N] = v[jj];
                      vmax = fmaxf(vmax, fabsf(v[jj]));
                    }
                  }
                  if (Lwrite) {
                    uint4 vh, vl;
                    split2(v[0] * s_out, v[1] * s_out, vh.x, vl.x);
                    split2(v[2] * s_out, v[3] * s_out, vh.y, vl.y);
                    split2(v[4] * s_out, v[5] * s_out, vh.z, vl.z);
                    split2(v[6] * s_out, v[7] * s_out, vh.w, vl.w);
                    const int64_t poff = (int64_t)((plane >> 1) * 4 + (plane & 1)) * g.slots;
                    uint4* act_hi = act_sample + poff;
                    uint4* act_lo = act_hi + 2 * g.slots;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                      if (q4 < rps[t].n) {
                        act_hi[rps[t].o[q4]] = vh;
                        act_lo[rps[t].o[q4]] = vl;
                      }
                    if (p.act_z) {  // interior slot only
                      uint4* z_hi = reinterpret_cast<uint4*>(p.act_z) + (int64_t)L.out_buf * p.buf_u4 + s * g.Ps + poff;
                      z_hi[rps[t].o[0]] = vh;
                      z_hi[2 * g.slots + rps[t].o[0]] = vl;
                    }
                  }
                }
                if (pix == 0 && cgp == 0) L.sig_out[s] = s_out;
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
              if (lane == 0 && vmax > 0.f) atomicMax(reinterpret_cast<unsigned*>(L.max_out + s), __float_as_uint(vmax));
            }
          };
          if constexpr (BWD) {
            run_bwd();
          } else {
            const bool fast_ok = p.epi_generic == 0;
            if (L.planar && fast_ok) run_fast(std::true_type{});
            else if (!Lres && !Lraw && !Lspin && fast_ok) run_fast(std::false_type{});
            else if (L.planar) run(std::true_type{});
            else run(std::false_type{});
          }
          // the next layer's TMA loads (either CTA) must see these stores: proxy fence by every writer, a named
          // barrier over the epilogue warps (orders all their stores before the signalling thread), then ONE thread
          // pays the cluster-scope release (it waits for the stores to be performed) while the other warps move on
          fence_proxy_async_global();
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (warp == 2 && lane == 0) {
            tc2_arrive_remote(ready_addr[j][0]);
            tc2_arrive_remote(ready_addr[j][1]);
          }
        }
    if (p.dbg && leader && warp == 2 && lane == 0) {
      unsigned long long* d = p.dbg + (size_t)pair_id * 16;
      d[8] = (unsigned long long)(clock64() - t_begin); d[9] = t_tfull; d[10] = t_drain;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc2_cluster_sync();  // the peer must not exit (or free TMEM) while the leader still reads its smem / TMEM
  if (warp == 1) tc2_tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// Jacobian on the tensor cores (variational.py:429-491): the backward-data layers are the SAME persistent CTA-pair
// kernel (mode 1 layers, transposed / flipped weight blobs); below are the seed of that tower and the per-sample
// weight gradients.
// ---------------------------------------------------------------------------------------------
// seed: raw d log psi / d x_last [ns, C, N] -> gradient operand buffer 0 (scaled per sample), max and scale arrays
__global__ void __launch_bounds__(256) tc_grad_seed_kernel(const float* __restrict__ dz, TcGeom g, int C, int Np,
                                                           __half* __restrict__ gbuf, __half* __restrict__ gzbuf,
                                                           float* __restrict__ gmax, float* __restrict__ gsig) {
  __shared__ float red[8];
  __shared__ float bc;
  const int64_t s = blockIdx.x;
  const int N = g.H * g.W, planes = Np >> 3;
  const float* d = dz + s * C * N;
  float m = 0.f;
  for (int e = threadIdx.x; e < C * N; e += blockDim.x) m = fmaxf(m, fabsf(d[e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bc = m;
    gmax[s] = m;
    gsig[s] = grad_scale(m);
  }
  __syncthreads();
  const float sig = grad_scale(bc);
  for (int e = threadIdx.x; e < planes * N; e += blockDim.x) {
    const int pix = e % N, plane = e / N;
    const PixSlots ps = pixel_slots(g, pix / g.W, pix % g.W);
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = plane * 8 + j;
      t[j] = (c < C) ? sig * d[c * N + pix] : 0.f;
    }
    store_plane(gbuf, g, plane, s * g.Ps, ps, t);
    if (gzbuf) {  // RASTER: the copy without halo slots (weight-gradient kernel)
      PixSlots p1 = ps;
      p1.n = 1;
      store_plane(gzbuf, g, plane, s * g.Ps, p1, t);
    }
  }
}

// Per-sample weight gradient of one convolution:
//   O[s, col0 + (o C + c) 9 + tap] = sum_pix dY[s, o, pix] * a[s, c, pix + tap]
// as a GEMM with M = o (128 rows, C used), N = c, K = pixels.  Both operands are the rasters the towers left behind,
// read MN-major: a 16-byte slot holds 8 channels of one pixel, 8 consecutive slots of a plane are a core matrix of
// the no-swizzle MN-major layout (SBO = plane pitch, LBO = 10 slots = the next 8-pixel segment), and the operand of
// tap (dy, dx) is the staged activation chunk addressed dy * row_pitch + dx slots further.  TMEM holds the nine tap
// accumulators of NPh in-channels (9 NPh <= 512 columns), so the in-channel planes are covered in `npass` passes;
// K runs over chunks of pixel rows (SEG) / of 16-slot steps (RASTER) staged by bulk copies (gradient slots +
// activation slots with one raster row of halo on either side).  On the RASTER layout K runs over ALL slots from the
// first to the last interior pixel, halo columns included, so the gradient operand must be ZERO there: the towers
// keep a second copy of every gradient operand without halo copies for this kernel (TcNetParams::act_z).
//   warp 0 producer, warp 1 MMA issuer, warps 2-13 epilogue (TMEM -> float64 / float32 Jacobian entries).
struct WgLayer {
  int g_buf, a_buf;   // gradient / activation operand buffers
  int64_t col0;       // first Jacobian column of this weight
};
struct WgParams {
  const __half* G;
  const __half* A;
  int64_t buf_halfs;
  TcGeom g;
  int C, Np, KS, nl;
  int nchunks, KK;       // chunks per sample, K steps (16 pixels) per chunk
  int runA, runB;        // slots per staged plane run of the gradient / activation chunk
  int chunk_stride;      // slots between consecutive chunks (KK K steps)
  int npass, PP;         // passes over the in-channel planes, planes per pass
  int stages;
  int pad_bytes;         // slack behind the last stage (the 128-row gradient operand reads 16 planes)
  float comp;            // expected-value correction of the truncating accumulator (tc_trunc_comp)
  const float* gsig;     // [buffers][ns]
  int64_t ns;
  void* out;
  int64_t ld;
  int out_f64;
  int vec_ok;            // rows of 8 channels x 9 taps are 32-byte (f64) / 16-byte (f32) aligned
  WgLayer layer[kTcMaxLayers];
};

// Epilogue: a TMEM lane (thread) holds one out-channel row, and consecutive Jacobian entries run over (in-channel,
// tap) of that row, so a warp storing straight from registers touches 32 cache lines per instruction; a transpose
// through shared memory with per-element global stores made the epilogue the longest phase (3.2 G instructions per
// 2048 samples), and one bulk copy per thread and row overran the copy engine with 288-byte requests
// (profiles/r2_wgrad_tc_epilogue_history.md).  So every thread converts 4 in-channels x 9 taps = 36 consecutive
// entries of ITS row into its line of the warp's [32 rows][36] box, and ONE tensor store per warp writes the box
// (cp.async.bulk.tensor, map = (in-channel * 9 + tap, out-channel, sample) of this layer's weight block; rows and
// columns beyond C are clipped by the map).  Two boxes per warp alternate.
constexpr int kWgThreads = 192;  // producer, MMA issuer, 4 epilogue warps (one per TMEM lane quarter)
constexpr int kWgEpiWarps = 4;
template <typename OutT>
struct WgLine {
  static constexpr int kPitch = 36 * (int)sizeof(OutT);  // dense box rows (the tensor store's shared-memory layout)
};
struct WgMaps {
  CUtensorMap m[kTcMaxLayers];  // per weight block: (in-channel * 9 + tap, out-channel, sample), box 36 x 32 x 1
};
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(kWgThreads, 1) resconv_wgrad_tc_kernel(const __grid_constant__ WgMaps maps,
                                                                        const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const TcGeom& g = p.g;
  const int planes = p.Np >> 3;
  const uint32_t runA16 = (uint32_t)p.runA, runB16 = (uint32_t)p.runB;  // slots per staged plane run
  const uint32_t offA_lo = (uint32_t)planes * runA16;
  const uint32_t offB_hi = 2u * (uint32_t)planes * runA16, offB_lo = offB_hi + (uint32_t)p.PP * runB16;
  const uint32_t stage16 = offB_hi + 2u * (uint32_t)p.PP * runB16;
  const uint32_t stage_bytes = stage16 * 16u;
  // [active quarter][2 boxes][32 lines], 128-byte aligned (tensor store source)
  unsigned char* stg = (unsigned char*)(((uintptr_t)(smem + (size_t)p.stages * stage_bytes + p.pad_bytes) + 127) & ~(uintptr_t)127);
  const int nact = (p.C + 31) >> 5;  // TMEM lane quarters that hold out-channel rows
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + (size_t)nact * 2 * 32 * WgLine<OutT>::kPitch);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kWgEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int NPh = p.PP * 8;
  const int64_t nunits = p.ns * p.nl;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int64_t s = u / p.nl;
        const WgLayer& L = p.layer[(int)(u - s * p.nl)];
        const __half* gsrc = p.G + (int64_t)L.g_buf * p.buf_halfs;
        const __half* asrc = p.A + (int64_t)L.a_buf * p.buf_halfs;
        for (int pass = 0; pass < p.npass; ++pass)
          for (int ch = 0; ch < p.nchunks; ++ch) {
            mbar_wait(empty + st, ph ^ 1);
            const int np_pass = min(p.PP, planes - pass * p.PP);  // planes of this pass
            const uint32_t bytes = (2u * (uint32_t)planes * runA16 + 2u * (uint32_t)np_pass * runB16) * 16u;
            mbar_expect_tx(full + st, bytes);
            unsigned char* dst = smem + (size_t)st * stage_bytes;
            const int64_t slotA = s * g.Ps + g.RP + (int64_t)ch * p.chunk_stride;  // raster row of the chunk's first pixel
            const int64_t slotB = slotA - g.RP;                                     // one row above it
            for (int hl = 0; hl < 2; ++hl)
              for (int pl = 0; pl < planes; ++pl) {
                const int64_t row = (pl >> 1) * 4 + hl * 2 + (pl & 1);
                bulk_g2s(dst + ((size_t)(hl * planes + pl) * runA16) * 16, gsrc + (row * g.slots + slotA) * 8, runA16 * 16u,
                         full + st);
              }
            for (int hl = 0; hl < 2; ++hl)
              for (int pp = 0; pp < np_pass; ++pp) {
                const int pl = pass * p.PP + pp;
                const int64_t row = (pl >> 1) * 4 + hl * 2 + (pl & 1);
                bulk_g2s(dst + ((size_t)offB_hi + (size_t)(hl * p.PP + pp) * runB16) * 16, asrc + (row * g.slots + slotB) * 8,
                         runB16 * 16u, full + st);
              }
            if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // D = F32, A = B = F16, both MN-major, M = 128, N = NPh
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(NPh >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo = (uint32_t)g.SB << 16;  // next 8-pixel segment
    const uint32_t a_hi_w = runA16 | (1u << 14), b_hi_w = runB16 | (1u << 14);
    const uint32_t smem16 = smem_u32(smem) >> 4;
    uint32_t st = 0, ph = 0, q = 0;
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x)
      for (int pass = 0; pass < p.npass; ++pass) {
        mbar_wait(tmem_empty, (q & 1) ^ 1);
        ++q;
        tc_fence_after();
        for (int ch = 0; ch < p.nchunks; ++ch) {
          mbar_wait(full + st, ph);
          tc_fence_after();
          const uint32_t base = smem16 + st * stage16;
          for (int kk = 0; kk < p.KK; ++kk) {
            const uint32_t a_h = (base + 1u + (uint32_t)kk * 2u * (uint32_t)g.SB) | lbo, a_l = a_h + offA_lo;
            const uint32_t b0 = (base + offB_hi + (uint32_t)kk * 2u * (uint32_t)g.SB) | lbo;
            const uint32_t acc0 = (ch == 0 && kk == 0) ? 0u : 1u;
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t b_h = b0 + (uint32_t)((tap / 3) * g.RP + (tap % 3)), b_l = b_h + (offB_lo - offB_hi);
                const uint32_t d = tmem_base + (uint32_t)(tap * NPh);
                umma_f16_split(d, a_h, a_hi_w, b_h, b_hi_w, idesc, acc0);
                umma_f16_split(d, a_l, a_hi_w, b_h, b_hi_w, idesc, 1u);
                umma_f16_split(d, a_h, a_hi_w, b_l, b_hi_w, idesc, 1u);
              }
              if (kk == p.KK - 1) {
                umma_commit(empty + st);
                if (ch == p.nchunks - 1) umma_commit(tmem_full);
              }
            }
            __syncwarp();
          }
          if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1; }
        }
      }
  } else {
    const int lq = warp & 3;
    const int o = lq * 32 + lane;
    const bool row_ok = o < p.C;
    unsigned char* box0 = stg + (size_t)(lq < nact ? lq : 0) * 2 * 32 * WgLine<OutT>::kPitch;
    unsigned char* box1 = box0 + (size_t)32 * WgLine<OutT>::kPitch;
    uint32_t q = 0;
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int64_t s = u / p.nl;
      const int li = (int)(u - s * p.nl);
      const WgLayer& L = p.layer[li];
      const float scale = p.comp / (__ldg(p.gsig + (int64_t)L.g_buf * p.ns + s) * kActScale);
      OutT* orow = reinterpret_cast<OutT*>(p.out) + s * p.ld + L.col0 + (int64_t)o * p.C * 9;
      for (int pass = 0; pass < p.npass; ++pass) {
        if (lane == 0) mbar_wait(tmem_full, q & 1);
        __syncwarp();
        ++q;
        tc_fence_after();
        int last = -1;  // the last in-channel group of this pass that exists (warp-uniform)
        if (lq < nact)
          for (int cg = 0; cg < p.PP; ++cg)
            if ((pass * p.PP + cg) * 8 < p.C) last = cg;
        for (int cg = 0; cg <= last; ++cg) {
          uint32_t r[9][8];
          __syncwarp();
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            tmem_ld8_nowait(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(tap * NPh + cg * 8), r[tap]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (cg == last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
          }
          const int c0 = (pass * p.PP + cg) * 8;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int nv = min(4, p.C - c0 - 4 * half);  // in-channels of this half that exist (warp-uniform)
            if (nv <= 0) break;
            if (p.vec_ok) {
              unsigned char* box = half ? box1 : box0;
              if (lane == 0) bulk_wait_read<1>();  // the store that last read this box (two groups ago) has drained it
              __syncwarp();
              unsigned char* line = box + (size_t)lane * WgLine<OutT>::kPitch;
#pragma unroll
              for (int e4 = 0; e4 < 9; ++e4) {  // entries e = jj * 9 + tap, four at a time
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int e = e4 * 4 + i;
                  v[i] = __uint_as_float(r[e % 9][4 * half + e / 9]) * scale;
                }
                if constexpr (sizeof(OutT) == 8) {
                  *reinterpret_cast<double2*>(line + e4 * 32) = make_double2((double)v[0], (double)v[1]);
                  *reinterpret_cast<double2*>(line + e4 * 32 + 16) = make_double2((double)v[2], (double)v[3]);
                } else {
                  *reinterpret_cast<float4*>(line + e4 * 16) = make_float4(v[0], v[1], v[2], v[3]);
                }
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&maps.m[li], box, (c0 + 4 * half) * 9, lq * 32, (int)s);
                bulk_commit();
              }
            } else if (row_ok) {
              OutT* dst = orow + (int64_t)(c0 + 4 * half) * 9;
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (jj < nv) {
#pragma unroll
                  for (int tap = 0; tap < 9; ++tap) dst[jj * 9 + tap] = (OutT)(__uint_as_float(r[tap][4 * half + jj]) * scale);
                }
            }
          }
        }
        if (last < 0) {  // nothing to drain (idle quarter / padded channels): the warp still takes part in the hand-over
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();  // the copy engine no longer reads this CTA's shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tc_encode_fn() {
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

static bool tc_geometry(int H, int W, int64_t ns, TcGeom& g, bool pair = false) {
  memset(&g, 0, sizeof(g));
  if (H < 2 || W < 2) return false;
  g.H = H; g.W = W; g.ns = ns;
  g.nseg = W / 8;
  const bool seg_ok = (W % 8 == 0) && ((H * g.nseg) % 16 == 0) && (H * g.nseg / 16 <= 2);
  if (seg_ok) {
    g.mode = 1; g.RP = g.nseg * 10; g.SB = 10; g.tps = H * g.nseg / 16;
  } else {
    g.mode = 0; g.nseg = 0; g.RP = W + 2; g.SB = 8;
    const int span = (H - 1) * g.RP + W;
    g.tps = (span + 127) / 128;
    if (g.tps > 2) return false;
  }
  g.Ps = (H + 2) * g.RP;
  g.pair = pair ? 1 : 0;
  const int ntile = pair ? 4 : 2;  // 128-row tiles per work item
  g.spi = ntile / g.tps;
  // slots read by one 128-row tile (all nine taps); a multiple of 16 so that half a plane run (TSh * 8 B, one TMA
  // box of the CTA-pair kernel) stays 128-byte aligned in shared memory
  g.TSh = (15 * g.SB + 7 + 2 * g.RP + 2 + 1 + 15) & ~15;
  int ts = 0;
  for (int q = 0; q < ntile; ++q) {
    const int sl = q / g.tps, tis = q % g.tps;
    const int need = sl * g.Ps + tis * 16 * g.SB + g.TSh;
    if (need > ts) ts = need;
  }
  if (ts < g.spi * g.Ps) ts = g.spi * g.Ps;
  g.TS = (ts + 1) & ~1;
  g.nitems = (ns + g.spi - 1) / g.spi;
  g.slots = g.nitems * g.spi * (int64_t)g.Ps + (g.TS - g.spi * g.Ps) + 16;
  return true;
}

// The float32 accumulator of tcgen05.mma rounds TOWARD ZERO at every accumulate, i.e. a sum of n
// instructions loses part of an ulp of the running sum n times, always in the direction of zero: a systematic shrink
// per accumulate relative to the result.  A NumPy model of a truncating accumulator (tools/trunc_bias_sim.py) gives
// 2.15 .. 2.33e-8 per accumulate for random-walk, coherent, biased and heavy-tailed sums alike (n = 18 .. 72); on the
// B200 the constant that minimises the error against a float64 evaluation is 1.75 .. 1.9e-8 for three network shapes
// (C = 32 / 88 / 128; profiles/r2_tc_trunc_comp_scan.md).  Left alone it is THE error of the towers -- 1e-6 per
// convolution at C = 88, in phase over all layers and samples: log psi 2.7e-6, Jacobian rows 1.2e-5 at config E;
// with the epilogues multiplying by the expected value 1 + 1.8e-8 n: 1e-7 and 1.1e-6.
// QTX_TC_TRUNC_COMP sets the constant in units of 1e-8 (0 = off).
static float tc_trunc_comp(int naccum) {
  double per = 1.8e-8;
  if (const char* e = getenv("QTX_TC_TRUNC_COMP")) per = atof(e) * 1e-8;
  return (float)(1.0 + per * naccum);
}

// tap-pair layout of a half-filled last K step (CTA-pair kernel): C % 16 in 1..8; QTX_TC_PAIRK=0 switches it off
static int tc_pair_k(int C, bool pair) {
  const char* e = getenv("QTX_TC_PAIRK");
  if (e && e[0] == '0') return 0;
  const int r = C & 15;
  return (pair && r >= 1 && r <= 8) ? 1 : 0;
}
// MMA instructions that feed the hi * hi accumulator of one tile and layer
static int tc_main_accumulates(int C, int pair_k) {
  const int KS = (C + 15) / 16;
  return pair_k ? 9 * (KS - 1) + 5 : 9 * KS;
}

static bool tc_disabled() {
  const char* e = getenv("QTX_RESCONV_TC");
  return e && e[0] == '0';
}

bool resconv_tc_supported(int C, int lx, int ly, int kh, int kw) {
  if (tc_disabled()) return false;
  if (kh != 3 || kw != 3 || C < 1 || C > 128) return false;
  TcGeom g;
  return tc_geometry(lx, ly, 1, g);
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static bool tc_use_pair() {
  const char* e = getenv("QTX_TC_2CTA");
  return e ? atoi(e) != 0 : true;
}

static void tc_sizes(int nblocks, int C, int lx, int ly, int64_t ns, TcGeom& g, int& Np, int& KS, size_t& blob_halfs,
                     size_t& act_bytes, size_t& wblob_bytes, size_t& resid_bytes) {
  tc_geometry(lx, ly, ns, g, tc_use_pair());
  Np = (C + 15) & ~15;
  KS = Np / 16;
  blob_halfs = (size_t)KS * 9 * 4 * Np * 8;
  wblob_bytes = align256((size_t)(2 * nblocks - 1) * blob_halfs * 2 + (size_t)(2 * nblocks - 1) * Np * 4);  // + padded biases
  act_bytes = align256((size_t)KS * 4 * g.slots * 16);
  resid_bytes = align256((size_t)ns * Np * lx * ly * 4);  // planar residual stream of the forward-only mode
}

static bool tc_bwd_disabled() {
  const char* e = getenv("QTX_RESCONV_TC_BWD");
  return e && e[0] == '0';
}

// the Jacobian runs on the tensor cores wherever the CTA-pair tower does (resconv_tc_backward)
bool resconv_tc_backward_supported(int C, int lx, int ly, int kh, int kw) {
  if (!resconv_tc_supported(C, lx, ly, kh, kw) || !tc_use_pair() || tc_bwd_disabled()) return false;
  TcGeom g;
  if (!tc_geometry(lx, ly, 1, g, true)) return false;
  if (g.mode != 1) {  // RASTER (dev knob: QTX_RESCONV_TC_BWD_RASTER=0 keeps these lattices on the CUDA cores)
    const char* e = getenv("QTX_RESCONV_TC_BWD_RASTER");
    if (e && e[0] == '0') return false;
  }
  return ((C + 15) & ~15) <= 128;
}

// workspace of the tensor-core Jacobian: every layer's operand rasters of both towers, both weight blob sets, the
// raw gradients [ns, C, N] of every convolution output, per-sample scales
struct TcBwdLayout {
  size_t buf_bytes;  // one operand buffer
  size_t op_off, wf_off, wb_off, g_off, gz_off, rg_off, gmax_off, gsig_off, wnorm_off, total;
  int nl, raster;
};
static TcBwdLayout tc_bwd_layout(int nblocks, int C, int lx, int ly, int64_t ns) {
  TcGeom g;
  int Np, KS;
  size_t bh, ab, wb, rb;
  tc_sizes(nblocks, C, lx, ly, ns, g, Np, KS, bh, ab, wb, rb);
  TcBwdLayout L{};
  L.nl = 2 * nblocks - 1;
  L.buf_bytes = (size_t)KS * 4 * g.slots * 16;  // exact: buffer b starts at tensor-map row b * KS * 4
  size_t off = 0;
  L.op_off = off; off += align256((size_t)L.nl * L.buf_bytes);
  L.wf_off = off; off += wb;
  L.wb_off = off; off += wb;
  L.g_off = off; off += align256((size_t)L.nl * L.buf_bytes);
  L.raster = g.mode == 0 ? 1 : 0;
  L.gz_off = off;  // RASTER: gradient operands without halo copies
  if (L.raster) off += align256((size_t)L.nl * L.buf_bytes);
  L.rg_off = off; off += (size_t)L.nl * align256((size_t)ns * C * lx * ly * 4);
  L.gmax_off = off; off += align256((size_t)(L.nl + 1) * ns * 4);
  L.gsig_off = off; off += align256((size_t)(L.nl + 1) * ns * 4);
  L.wnorm_off = off; off += align256((size_t)L.nl * 4);
  L.total = off + 512;
  return L;
}

size_t resconv_tc_workspace(int64_t ns, int nblocks, int C, int lx, int ly, int grad) {
  if (grad && resconv_tc_backward_supported(C, lx, ly, 3, 3)) return tc_bwd_layout(nblocks, C, lx, ly, ns).total;
  TcGeom g;
  int Np, KS;
  size_t bh, ab, wb, rb;
  tc_sizes(nblocks, C, lx, ly, ns, g, Np, KS, bh, ab, wb, rb);
  return ab + wb + rb + 512;
}

// launch the persistent tower kernel over the layers of `np` (operand buffers: `nbuf` raster sets behind np.act)
static int tc_launch_tower(TcNetParams& np, int nbuf, int nl, cudaStream_t st) {
  const TcGeom& g = np.g;
  const int Np = np.Np, KS = np.KS;
  __half* act = np.act;
  const __half* wblob = np.wblob;
  // shared memory: activation ring + weight ring + barriers
  const size_t act_stage = g.pair ? (size_t)2 * 4 * g.TSh * 16 : (size_t)4 * g.TS * 16;
  const size_t w_stage = g.pair ? (size_t)3 * 4 * (Np / 2) * 16 : (size_t)3 * 4 * Np * 16;
  int act_stages = 3, w_stages = g.pair ? 9 : 6;
  const size_t cap = 227 * 1024 - 1024;
  while (act_stages > 2 && act_stages * act_stage + w_stages * w_stage > cap) --act_stages;
  while (w_stages > 2 && act_stages * act_stage + w_stages * w_stage > cap) --w_stages;
  QTX_REQUIRE(act_stages * act_stage + w_stages * w_stage <= cap, QTX_ERR_UNSUPPORTED, "resconv_tc: tile does not fit");
  if (const char* e = getenv("QTX_TC_WSTAGES")) { int v = atoi(e); if (v >= 2 && act_stages * act_stage + v * w_stage <= cap) w_stages = v; }
  np.act_stages = act_stages; np.w_stages = w_stages;
  const size_t smem = act_stages * act_stage + w_stages * w_stage + (2 * act_stages + 2 * w_stages + 4) * 8 + 16 + 128;
  const int PL = (Np / 8 + kColGroups - 1) / kColGroups;
  int per_launch = nl;
  if (const char* e = getenv("QTX_TC_LAYERS_PER_LAUNCH")) { int v = atoi(e); if (v >= 1) per_launch = v; }
  static unsigned long long* dbg_buf = nullptr;
  const bool dbg = getenv("QTX_TC_DEBUG") != nullptr;
  if (dbg && !dbg_buf) cudaMalloc(&dbg_buf, 256 * 16 * sizeof(unsigned long long));
  np.dbg = dbg ? dbg_buf : nullptr;
    auto report = [&](int units) {
    static unsigned long long h[256 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    double a[16] = {0};
    for (int b2 = 0; b2 < units; ++b2)
      for (int i = 0; i < 16; ++i) a[i] += (double)h[b2 * 16 + i] / units;
    fprintf(stderr,
            "[tc dbg%s] layers %d..%d items/unit %.1f | producer total %.0f wait: act_ready %.0f act_empty %.0f w_empty %.0f | "
            "mma total %.0f wait: tmem_empty %.0f act_full %.0f w_full %.0f | epi total %.0f wait tmem_full %.0f drain %.0f\n",
            g.pair ? " 2cta" : "", np.layer0, np.layer1, (double)g.nitems / units, a[0], a[1], a[2], a[3], a[4], a[5], a[6],
            a[7], a[8], a[9], a[10]);
  };
  if (g.pair) {
    EncodeTiledFn encode = tc_encode_fn();
    QTX_REQUIRE(encode != nullptr, QTX_ERR_CUDA, "resconv_tc: cuTensorMapEncodeTiled is unavailable");
    CUtensorMap tmA, tmW;
    {
      // operand rasters as 8-byte elements: (2 * slot, plane = kstep * 4 + hi|lo * 2 + p2); box = half a plane run
      cuuint64_t gdim[2] = {(cuuint64_t)g.slots * 2, (cuuint64_t)KS * 4 * (cuuint64_t)nbuf};
      cuuint64_t gstride[1] = {(cuuint64_t)g.slots * 16};
      cuuint32_t box[2] = {(cuuint32_t)g.TSh, 1};
      cuuint32_t estr[2] = {1, 1};
      CUresult cr = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, act, gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      QTX_REQUIRE(cr == CUDA_SUCCESS, QTX_ERR_CUDA, "resconv_tc: cuTensorMapEncodeTiled (activations) failed (%d)", (int)cr);
    }
    {
      // weight blobs: rows of 4 * Np/2 8-byte elements; one (layer, kstep, dy, half) stage = 6 rows
      const cuuint64_t rowel = (cuuint64_t)4 * (Np / 2);
      cuuint64_t gdim[2] = {rowel, (cuuint64_t)nl * KS * 36};
      cuuint64_t gstride[1] = {rowel * 8};
      cuuint32_t box[2] = {(cuuint32_t)rowel, 6};
      cuuint32_t estr[2] = {1, 1};
      CUresult cr = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<__half*>(wblob), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      QTX_REQUIRE(cr == CUDA_SUCCESS, QTX_ERR_CUDA, "resconv_tc: cuTensorMapEncodeTiled (weights) failed (%d)", (int)cr);
    }
    void (*kern)(CUtensorMap, CUtensorMap, TcNetParams) = nullptr;
    const bool bwd = np.layer[0].mode == 1;
    switch (PL) {
      case 1: kern = bwd ? resconv_tc2_kernel<1, true> : resconv_tc2_kernel<1, false>; break;
      case 2: kern = bwd ? resconv_tc2_kernel<2, true> : resconv_tc2_kernel<2, false>; break;
      case 3: kern = bwd ? resconv_tc2_kernel<3, true> : resconv_tc2_kernel<3, false>; break;
      case 4: kern = bwd ? resconv_tc2_kernel<4, true> : resconv_tc2_kernel<4, false>; break;
      case 5: kern = bwd ? resconv_tc2_kernel<5, true> : resconv_tc2_kernel<5, false>; break;
      default: kern = bwd ? resconv_tc2_kernel<6, true> : resconv_tc2_kernel<6, false>; break;
    }
    QTX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int npairs = num_sms() / 2;
    if ((int64_t)npairs > g.nitems) npairs = (int)g.nitems;
    for (int l0 = 0; l0 < nl; l0 += per_launch) {
      np.layer0 = l0;
      np.layer1 = l0 + per_launch < nl ? l0 + per_launch : nl;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(2 * npairs));
      cfg.blockDim = dim3(kTcThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      QTX_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, np));
      count_launch();
      if (dbg) report(npairs);
    }
    return QTX_OK;
  }
  void (*kern)(TcNetParams) = nullptr;
  switch (PL) {
    case 1: kern = resconv_tc_kernel<1>; break;
    case 2: kern = resconv_tc_kernel<2>; break;
    case 3: kern = resconv_tc_kernel<3>; break;
    case 4: kern = resconv_tc_kernel<4>; break;
    case 5: kern = resconv_tc_kernel<5>; break;
    default: kern = resconv_tc_kernel<6>; break;
  }
  QTX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = num_sms();
  if ((int64_t)grid > g.nitems) grid = (int)g.nitems;
  for (int l0 = 0; l0 < nl; l0 += per_launch) {
    np.layer0 = l0;
    np.layer1 = l0 + per_launch < nl ? l0 + per_launch : nl;
    kern<<<grid, kTcThreads, smem, st>>>(np);
    QTX_LAUNCH_CHECK();
    if (dbg) report(grid);
  }
  return QTX_OK;
}

// Forward through the tower.
//   save_all == 0: *x_final receives the residual stream x_nblocks in the PLANAR layout [ns, Np/8, N, 8]
//                  (it lives in the tensor-core workspace), X / Hs are not touched
//   save_all == 1: X[i] = X + i*act holds x_{i+1}, Hs[i] = Hs + i*act the conv1 pre-activation of block i,
//                  both [ns, C, N] (what the backward kernels of resconv.cu read); *x_final = X[nblocks-1]
int resconv_tc_forward(int nblocks, int C, int lx, int ly, const float* params, const int8_t* spins, int64_t ns,
                       float* X, float* Hs, int save_all, void* ws, size_t ws_bytes, const float** x_final,
                       int* x_final_planes, const long long* ns_dev, cudaStream_t st) {
  TcGeom g;
  int Np, KS;
  size_t blob_halfs, act_bytes, wblob_bytes, resid_bytes;
  tc_sizes(nblocks, C, lx, ly, ns, g, Np, KS, blob_halfs, act_bytes, wblob_bytes, resid_bytes);
  QTX_REQUIRE(2 * nblocks - 1 <= kTcMaxLayers, QTX_ERR_UNSUPPORTED, "resconv_tc: too many blocks");
  unsigned char* base = (unsigned char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  // save_all == 2: layout of the tensor-core Jacobian (every layer keeps its operand buffer)
  const bool keep_ops = save_all == 2;
  const TcBwdLayout bl = keep_ops ? tc_bwd_layout(nblocks, C, lx, ly, ns) : TcBwdLayout{};
  if (keep_ops) {
    QTX_REQUIRE(g.pair, QTX_ERR_UNSUPPORTED, "resconv_tc: the tensor-core Jacobian needs the CTA-pair kernel");
    QTX_REQUIRE(ws_bytes >= bl.total - 256, QTX_ERR_INVALID, "resconv_tc: workspace too small");
  } else {
    QTX_REQUIRE(ws_bytes >= act_bytes + wblob_bytes + resid_bytes + 256, QTX_ERR_INVALID, "resconv_tc: workspace too small");
  }
  __half* act = reinterpret_cast<__half*>(base + (keep_ops ? bl.op_off : 0));
  __half* wblob = reinterpret_cast<__half*>(base + (keep_ops ? bl.wf_off : act_bytes));
  float* resid = keep_ops ? nullptr : reinterpret_cast<float*>(base + act_bytes + wblob_bytes);
  float* bias_pad = reinterpret_cast<float*>(wblob + (size_t)(2 * nblocks - 1) * blob_halfs);
  const int N = lx * ly;
  const int64_t actsz = ns * C * N;

  // parameter offsets (ravel_pytree order, see resconv.cu)
  int64_t off = 0;
  int64_t w1[64], b1[64], w2[64], b2[64];
  for (int i = 0; i < nblocks; ++i) {
    w1[i] = off; off += (int64_t)C * (i == 0 ? 1 : C) * 9;
    b1[i] = off; off += C;
    w2[i] = off; off += (int64_t)C * C * 9;
    if (i == nblocks - 1) b2[i] = -1;
    else { b2[i] = off; off += C; }
  }

  // tensor-core layers: conv2_0, then (conv1_i, conv2_i) for i >= 1
  TcPrepParams pp{};
  pp.params = params; pp.wblob = wblob; pp.C = C; pp.Np = Np; pp.KS = KS; pp.pair = g.pair;
  pp.pair_k = tc_pair_k(C, g.pair != 0);
  TcNetParams np{};
  np.act = act; np.wblob = wblob; np.g = g; np.C = C; np.Np = Np; np.KS = KS;
  np.buf_u4 = (int64_t)KS * 4 * g.slots;
  np.pair_k = pp.pair_k;
  {
    const char* e = getenv("QTX_TC_PRECISE_GELU");  // dev knob: accurate exp / division in gelu (no measurable effect)
    np.precise = e ? atoi(e) : 0;
  }
  np.out_scale = kOutScale * tc_trunc_comp(tc_main_accumulates(C, np.pair_k));  // hi * hi accumulator
  {
    const char* e = getenv("QTX_TC_EPI_GENERIC");
    np.epi_generic = (e && e[0] == '1') ? 1 : 0;
    e = getenv("QTX_TC_KEEP_PAD");
    np.keep_pad = (e && e[0] == '1') ? 1 : 0;
  }
  int nl = 0;
  for (int i = 0; i < nblocks; ++i) {
    if (i > 0) {
      TcLayer& L = np.layer[nl];
      pp.w_off[nl] = w1[i]; pp.blob_off[nl] = (int64_t)nl * blob_halfs;
      L.wblob_off = pp.blob_off[nl];
      pp.b_off[nl] = b1[i];
      L.bias = bias_pad + (size_t)nl * Np; L.res = nullptr; L.res_spin = nullptr;
      L.raw_out = save_all ? Hs + (int64_t)i * actsz : nullptr;
      L.out_alpha = 1.0f; L.write_act = 1; L.planar = 0;
      L.in_buf = keep_ops ? nl : 0; L.out_buf = keep_ops ? nl + 1 : 0;
      ++nl;
    }
    TcLayer& L = np.layer[nl];
    L.in_buf = keep_ops ? nl : 0; L.out_buf = keep_ops ? nl + 1 : 0;
    pp.w_off[nl] = w2[i]; pp.blob_off[nl] = (int64_t)nl * blob_halfs;
    L.wblob_off = pp.blob_off[nl];
    pp.b_off[nl] = b2[i];
    L.bias = bias_pad + (size_t)nl * Np;
    L.res_spin = (i == 0) ? spins : nullptr;
    if (save_all) {
      L.res = (i == 0) ? nullptr : X + (int64_t)(i - 1) * actsz;
      L.raw_out = X + (int64_t)i * actsz;
      L.planar = 0;
    } else {
      L.res = (i == 0) ? nullptr : resid;
      L.raw_out = resid;
      L.planar = 1;
    }
    L.out_alpha = (float)(1.0 / sqrt((double)(i + 2)));
    L.write_act = (i < nblocks - 1) ? 1 : 0;
    ++nl;
  }
  pp.nconv = nl;
  pp.bias_pad = bias_pad;
  if (keep_ops && g.mode == 0)
    // RASTER: the K axis of the weight-gradient GEMM runs a few slots past the last interior pixel of a sample; the
    // gradient is zero there, so the activation slots it meets only have to be finite -- also in the guard behind the
    // last sample, which no kernel writes
    QTX_CUDA(cudaMemsetAsync(act, 0, (size_t)nl * bl.buf_bytes, st));
  if (x_final) *x_final = save_all ? X + (int64_t)(nblocks - 1) * actsz : resid;
  if (x_final_planes) *x_final_planes = save_all ? 0 : (Np >> 3);
  {
    const int n = 9 * KS * 16 * Np;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)nl);
    tc_weight_prep_kernel<<<grid, 256, 0, st>>>(pp);
    QTX_LAUNCH_CHECK();
  }
  {
    const int64_t total = ns * N;
    unsigned gsz = (unsigned)((total + 255) / 256);
    if (gsz > 16u * num_sms()) gsz = 16u * num_sms();
    tc_first_layer_kernel<<<gsz, 256, (size_t)Np * 10 * sizeof(float), st>>>(spins, params + w1[0], params + b1[0], g, C,
                                                                             Np, act, save_all ? Hs : nullptr, ns_dev, np.precise);
    QTX_LAUNCH_CHECK();
  }
  np.ns_dev = ns_dev;
  return tc_launch_tower(np, keep_ops ? nl : 1, nl, st);
}

// Jacobian rows of the tensor-core layers (variational.py:429-491).  Requires resconv_tc_forward(save_all = 2) on the
// same workspace (operand buffers of every layer, raw X / Hs) and the seed d log psi / d x_last in `seed` [ns, C, N].
//   raw_grad[k], k = 0 .. 2 nblocks - 1: raw gradient w.r.t. the output of  conv2_{nb-1}, conv1_{nb-1}, conv2_{nb-2}, ...
//   (raw_grad[0] = seed); the caller turns them into bias gradients and the first layer's weight gradient.
// The weight gradients of all convolutions with C input channels are written here.
int resconv_tc_backward(int nblocks, int C, int lx, int ly, const float* params, int64_t ns, const float* X,
                        const float* Hs, const float* seed, void* out, int out_f64, int64_t ld, void* ws, size_t ws_bytes,
                        const float** raw_grad, cudaStream_t st) {
  TcGeom g;
  int Np, KS;
  size_t blob_halfs, act_bytes, wblob_bytes, resid_bytes;
  tc_sizes(nblocks, C, lx, ly, ns, g, Np, KS, blob_halfs, act_bytes, wblob_bytes, resid_bytes);
  QTX_REQUIRE(g.pair, QTX_ERR_UNSUPPORTED, "resconv_tc_backward: needs the CTA-pair kernel");
  const TcBwdLayout bl = tc_bwd_layout(nblocks, C, lx, ly, ns);
  QTX_REQUIRE(ws_bytes >= bl.total - 256, QTX_ERR_INVALID, "resconv_tc_backward: workspace too small");
  unsigned char* base = (unsigned char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  const int nl = bl.nl, N = lx * ly, planes = Np >> 3;
  const int64_t actsz = ns * C * N;
  __half* OP = reinterpret_cast<__half*>(base + bl.op_off);
  __half* G = reinterpret_cast<__half*>(base + bl.g_off);
  __half* Gz = bl.raster ? reinterpret_cast<__half*>(base + bl.gz_off) : nullptr;
  __half* wblob = reinterpret_cast<__half*>(base + bl.wb_off);
  float* bias_pad = reinterpret_cast<float*>(wblob + (size_t)nl * blob_halfs);
  float* gmax = reinterpret_cast<float*>(base + bl.gmax_off);
  float* gsig = reinterpret_cast<float*>(base + bl.gsig_off);
  float* wnorm = reinterpret_cast<float*>(base + bl.wnorm_off);
  const size_t rg_stride = align256((size_t)actsz * 4);
  auto RG = [&](int k) -> float* { return k == 0 ? const_cast<float*>(seed) : reinterpret_cast<float*>(base + bl.rg_off + (size_t)(k - 1) * rg_stride); };
  for (int k = 0; k <= nl; ++k) raw_grad[k] = RG(k);

  int64_t off = 0;
  int64_t w1[64], w2[64];
  for (int i = 0; i < nblocks; ++i) {
    w1[i] = off; off += (int64_t)C * (i == 0 ? 1 : C) * 9 + C;
    w2[i] = off; off += (int64_t)C * C * 9 + (i == nblocks - 1 ? 0 : C);
  }

  // ---- backward-data tower: layer d reads gradient buffer d, writes d + 1 ----
  TcPrepParams pp{};
  pp.params = params; pp.wblob = wblob; pp.C = C; pp.Np = Np; pp.KS = KS; pp.pair = g.pair;
  pp.transpose = 1; pp.wnorm = wnorm; pp.bias_pad = bias_pad;
  pp.pair_k = tc_pair_k(C, g.pair != 0);
  TcNetParams np{};
  np.act = G; np.wblob = wblob; np.g = g; np.C = C; np.Np = Np; np.KS = KS;
  np.buf_u4 = (int64_t)KS * 4 * g.slots;
  {
    const char* e = getenv("QTX_TC_PRECISE_GELU");
    np.precise = e ? atoi(e) : 0;
  }
  np.pair_k = pp.pair_k;
  {
    const char* e = getenv("QTX_TC_KEEP_PAD");
    np.keep_pad = (e && e[0] == '1') ? 1 : 0;
  }
  np.out_scale = kOutScale * tc_trunc_comp(tc_main_accumulates(C, np.pair_k));
  np.act_z = Gz;
  WgParams wp{};
  int d = 0, nw = 0;
  for (int i = nblocks - 1; i >= 0; --i) {
    for (int which = 0; which < 2; ++which) {  // 0: through conv2_i, 1: through conv1_i
      if (which == 1 && i == 0) break;
      TcLayer& L = np.layer[d];
      pp.w_off[d] = which == 0 ? w2[i] : w1[i];
      pp.b_off[d] = -1;
      pp.blob_off[d] = (int64_t)d * blob_halfs;
      L.wblob_off = pp.blob_off[d];
      L.bias = bias_pad + (size_t)d * Np;  // zeros
      L.res_spin = nullptr; L.out_alpha = 1.0f; L.planar = 0;
      L.mode = 1;
      L.in_buf = d; L.out_buf = d + 1;
      L.sig_in = gsig + (int64_t)d * ns; L.max_in = gmax + (int64_t)d * ns;
      L.sig_out = gsig + (int64_t)(d + 1) * ns; L.max_out = gmax + (int64_t)(d + 1) * ns;
      L.wnorm = wnorm + d;
      L.raw_out = RG(d + 1);
      if (which == 0) {
        L.mul = Hs + (int64_t)i * actsz; L.mul_alpha = 1.0f; L.res = nullptr; L.max_res = nullptr;
        L.write_act = i > 0 ? 1 : 0;
      } else {
        L.mul = X + (int64_t)(i - 1) * actsz; L.mul_alpha = (float)(1.0 / sqrt((double)(i + 1)));
        L.res = RG(d - 1); L.max_res = gmax + (int64_t)(d - 1) * ns;
        L.write_act = 1;
      }
      // weight gradient of the convolution whose OUTPUT gradient is buffer d: conv2_i reads operand 2 i, conv1_i 2 i - 1
      WgLayer& W = wp.layer[nw++];
      W.g_buf = d;
      W.a_buf = which == 0 ? 2 * i : 2 * i - 1;
      W.col0 = which == 0 ? w2[i] : w1[i];
      ++d;
    }
  }
  // conv1_0 (one input channel) is left to the caller; its output gradient is raw_grad[nl]
  QTX_REQUIRE(d == nl, QTX_ERR_INVALID, "resconv_tc_backward: layer count");
  pp.nconv = nl;
  QTX_CUDA(cudaMemsetAsync(gmax, 0, (size_t)(nl + 1) * ns * 4, st));
  if (Gz)  // halo slots and the guard behind the last sample stay zero (K of the weight-gradient GEMM crosses them)
    QTX_CUDA(cudaMemsetAsync(Gz, 0, (size_t)nl * bl.buf_bytes, st));
  {
    const int n = 9 * KS * 16 * Np;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)nl);
    tc_weight_prep_kernel<<<grid, 256, 0, st>>>(pp);
    QTX_LAUNCH_CHECK();
    tc_wnorm_kernel<<<nl, 256, 0, st>>>(pp);
    QTX_LAUNCH_CHECK();
  }
  tc_grad_seed_kernel<<<(unsigned)ns, 256, 0, st>>>(seed, g, C, Np, G, Gz, gmax, gsig);
  QTX_LAUNCH_CHECK();
  {
    int rc = tc_launch_tower(np, nl, nl, st);
    if (rc) return rc;
  }

  // ---- weight gradients ----
  if (const char* e = getenv("QTX_TC_WGRAD")) {  // dev knob: 0 = the caller computes them from raw_grad (CUDA cores)
    if (e[0] == '0') return QTX_OK;
  }
  wp.G = Gz ? Gz : G; wp.A = OP; wp.buf_halfs = (int64_t)KS * 4 * g.slots * 8; wp.g = g; wp.C = C; wp.Np = Np; wp.KS = KS;
  wp.nl = nw; wp.gsig = gsig; wp.ns = ns; wp.out = out; wp.ld = ld; wp.out_f64 = out_f64;
  wp.PP = planes >= 6 ? 6 : ((planes + 1) & ~1);
  wp.npass = (planes + wp.PP - 1) / wp.PP;
  const size_t line_bytes = (size_t)((C + 31) / 32) * 2 * 32 * (out_f64 ? WgLine<double>::kPitch : WgLine<float>::kPitch);
  const size_t cap = 227 * 1024 - 1024 - line_bytes - 128;  // minus the epilogue's output boxes
  // K axis: SEG = interior 8-pixel segments (two per step), RASTER = all slots from the first to the last interior
  // pixel, padded to whole steps of 16 (the padding lies in the zeroed halo of the gradient copy)
  const int kstride = 2 * g.SB;  // slots per K step
  const int Ktot = g.mode ? g.H * g.nseg / 2 : ((g.H - 1) * g.RP + g.W + 15) / 16;
  int KK = 0;
  size_t stage = 0, pad = 0;
  for (int kk = Ktot; kk >= 1; --kk) {
    if (Ktot % kk) continue;
    const int runA = kk * kstride + (g.mode ? 0 : 2), runB = kk * kstride + 2 * g.RP + 2;
    const size_t sb = ((size_t)2 * planes * runA + (size_t)2 * wp.PP * runB) * 16;
    // M = 128 reads 16 planes from each gradient half: the rows beyond Np alias whatever follows in shared memory
    // (results of those rows are never read); behind the last stage that needs slack
    const size_t reach = (size_t)(planes + 16) * runA * 16;
    const size_t pd = reach > sb ? ((reach - sb + 127) & ~(size_t)127) : 0;
    if (3 * sb + pd + 256 <= cap || kk == 1) { KK = kk; stage = sb; pad = pd; wp.runA = runA; wp.runB = runB; break; }
  }
  QTX_REQUIRE(KK > 0 && 2 * stage + pad + 256 <= cap, QTX_ERR_UNSUPPORTED, "resconv_tc_backward: no K chunk fits");
  wp.KK = KK; wp.nchunks = Ktot / KK; wp.chunk_stride = KK * kstride;
  wp.comp = tc_trunc_comp(3 * Ktot);  // one accumulator: three products per 16-pixel step
  int stages = (int)((cap - 256 - pad) / stage);
  if (stages > 6) stages = 6;
  wp.stages = stages;
  wp.pad_bytes = (int)pad;
  {
    bool ok = true;
    const int64_t al = out_f64 ? 2 : 4;
    if (C % al || ld % al || ((uintptr_t)out & 15)) ok = false;
    for (int i = 0; i < nw; ++i)
      if (wp.layer[i].col0 % al) ok = false;
    wp.vec_ok = ok ? 1 : 0;
  }
  WgMaps wmaps;
  memset(&wmaps, 0, sizeof(wmaps));
  if (wp.vec_ok) {
    EncodeTiledFn encode = tc_encode_fn();
    QTX_REQUIRE(encode != nullptr, QTX_ERR_CUDA, "resconv_tc_backward: cuTensorMapEncodeTiled is unavailable");
    const size_t esz = out_f64 ? 8 : 4;
    for (int i = 0; i < nw; ++i) {
      cuuint64_t gdim[3] = {(cuuint64_t)C * 9, (cuuint64_t)C, (cuuint64_t)ns};
      cuuint64_t gstride[2] = {(cuuint64_t)C * 9 * esz, (cuuint64_t)ld * esz};
      cuuint32_t box[3] = {36, 32, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult cr = encode(&wmaps.m[i], out_f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           (unsigned char*)out + (size_t)wp.layer[i].col0 * esz, gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      QTX_REQUIRE(cr == CUDA_SUCCESS, QTX_ERR_CUDA, "resconv_tc_backward: cuTensorMapEncodeTiled (Jacobian) failed (%d)", (int)cr);
    }
  }
  const size_t smem = (size_t)stages * stage + pad + line_bytes + 128 + (size_t)(2 * stages + 2) * 8 + 16 + 128;
  int64_t nunits = ns * nw;
  int grid = num_sms();
  if ((int64_t)grid > nunits) grid = (int)nunits;
  if (out_f64) {
    QTX_CUDA(cudaFuncSetAttribute(resconv_wgrad_tc_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resconv_wgrad_tc_kernel<double><<<grid, kWgThreads, smem, st>>>(wmaps, wp);
  } else {
    QTX_CUDA(cudaFuncSetAttribute(resconv_wgrad_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resconv_wgrad_tc_kernel<float><<<grid, kWgThreads, smem, st>>>(wmaps, wp);
  }
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

}  // namespace qtx
