// RBM_Dense kernels: direct forward, persistent Metropolis sweep, fused local energies,
// ref_forward, Jacobian (optionally centred/scaled) and its column mean.
//
// Design (B200): one warp owns one Markov chain / sample.  The M hidden pre-activations
// theta live in registers (hidden unit i = r*32 + lane), the transposed weight matrix W^T
// [N, M] is staged once per CTA in shared memory when it fits (160 KB for the 10x10, alpha=4
// benchmark), so a proposal costs nflips conflict-free shared-memory rows, M/32 log-cosh
// evaluations per lane and one shuffle reduction -- no HBM traffic inside the sweep.
#include "common.cuh"

namespace qtx {

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_w_kernel(const T* __restrict__ W, T* __restrict__ Wt, int N, int M) {
  __shared__ T tile[32][33];
  int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int i = i0 + r, j = j0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < M && j < N) ? W[(size_t)i * N + j] : T(0);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int j = j0 + r, i = i0 + threadIdx.x;
    if (i < M && j < N) Wt[(size_t)j * M + i] = tile[threadIdx.x][r];
  }
}

template <typename T>
static int transpose_w(const void* W, void* Wt, int N, int M, cudaStream_t st) {
  dim3 grid((N + 31) / 32, (M + 31) / 32), block(32, 8);
  transpose_w_kernel<T><<<grid, block, 0, st>>>((const T*)W, (T*)Wt, N, M);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

// theta = W s + b (sum over sites first, bias last, as eqx.nn.Linear does), warp per sample
template <typename T>
__global__ void rbm_forward_kernel(const T* __restrict__ W, const T* __restrict__ b, int N, int M,
                                   const int8_t* __restrict__ spins, int64_t ns, T* __restrict__ theta_out,
                                   double* __restrict__ logabs_out) {
  int lane = threadIdx.x & 31;
  int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= ns) return;
  const int8_t* sp = spins + s * N;
  T la = 0;
  for (int i = 0; i < M; ++i) {
    T acc = 0;
    for (int j = lane; j < N; j += 32) acc += W[(size_t)i * N + j] * (T)sp[j];
    acc = warp_sum(acc) + b[i];
    if (theta_out && lane == 0) theta_out[s * M + i] = acc;
    la += lncosh(acc);
  }
  if (logabs_out && lane == 0) logabs_out[s] = (double)la;
}

// ---------------------------------------------------------------------------------------------
// persistent sweep
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SweepParams {
  const T* Wt;  // [N, M]
  const T* b;   // [M]
  int N, M;
  int8_t* spins;
  int64_t ns;
  int nsweeps, kind;
  const int32_t* nbr;
  int max_nb, hop;
  double reweight;
  const int32_t* inj_pos;
  const int32_t* inj_slot;
  const double* inj_u;
  uint32_t seed_lo, seed_hi;
  uint64_t step0, chain0;
  double* logabs_out;
  double* logabs_chain_out;
  int32_t* naccept_out;
  uint8_t* accept_log;
};

template <typename T>
__device__ __forceinline__ void stage_w(T* dst, const T* __restrict__ src, size_t n) {
  // 16-byte vector copy global -> shared (both 16B aligned by construction)
  const size_t nvec = (n * sizeof(T)) / 16;
  const int4* s4 = reinterpret_cast<const int4*>(src);
  int4* d4 = reinterpret_cast<int4*>(dst);
  for (size_t k = threadIdx.x; k < nvec; k += blockDim.x) d4[k] = __ldg(s4 + k);
  for (size_t k = nvec * (16 / sizeof(T)) + threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

__device__ __forceinline__ int spin_from_words(uint32_t myword, int j) {
  uint32_t w = __shfl_sync(FULL, myword, j >> 5);
  return ((w >> (j & 31)) & 1u) ? 1 : -1;
}

template <typename T, int RMAX, bool WSMEM>
__global__ void __launch_bounds__(RMAX <= 8 ? 1024 : (sizeof(T) == 4 && RMAX <= 16 ? 896 : 512), 1)
    rbm_sweep_kernel(SweepParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N, M = p.M;
  const T* Wt = p.Wt;
  if (WSMEM) {
    T* Ws = reinterpret_cast<T*>(smem_raw);
    stage_w(Ws, p.Wt, (size_t)N * M);
    __syncthreads();
    Wt = Ws;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * wpb;
  const int NW = (N + 31) >> 5;

  for (int64_t chain = (int64_t)blockIdx.x * wpb + warp; chain < p.ns; chain += nwarps) {
    int8_t* sp = p.spins + chain * N;
    // spins -> bit words (bit = 1 for spin up); lane w keeps word w
    uint32_t myword = 0;
    for (int w = 0; w < NW; ++w) {
      int j = w * 32 + lane;
      uint32_t bal = __ballot_sync(FULL, j < N && sp[j] > 0);
      if (lane == w) myword = bal;
    }
    // theta = W s + b
    T th[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) th[r] = 0;
    for (int j = 0; j < N; ++j) {
      T sj = (T)spin_from_words(myword, j);
      const T* col = Wt + (size_t)j * M;
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        int i = r * 32 + lane;
        if (i < M) th[r] += col[i] * sj;
      }
    }
    T lsum = 0;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      int i = r * 32 + lane;
      if (i < M) {
        th[r] += p.b[i];
        lsum += lncosh(th[r]);
      }
    }
    double la = (double)warp_sum(lsum);
    int nacc = 0;

    // Acceptance is tested in the log domain: |psi'/psi|^n > 1 - u  <=>  n (la' - la) > log(1 - u).
    // In Philox mode lane l draws the randoms of step (t & ~31) + l, so one float64 log per lane serves
    // 32 steps of the chain (instead of one float64 exp per step).
    uint32_t q0 = 0, q1 = 0;  // Philox outputs of step (t & ~31) + lane
    double qthr = 0.0;        // log(1 - u) of that step
    for (int t = 0; t < p.nsweeps; ++t) {
      int pos, slot = 0;
      double thr;
      if (p.inj_u) {
        size_t o = (size_t)t * p.ns + chain;
        pos = p.inj_pos[o];
        if (p.kind == QTX_SPIN_EXCHANGE) slot = p.inj_slot[o];
        thr = log1p(-p.inj_u[o]);
      } else {
        if ((t & 31) == 0) {
          uint64_t step = p.step0 + (uint64_t)(t + lane);
          uint32_t q2 = (uint32_t)(step >> 32), q3 = 0;
          q0 = (uint32_t)(p.chain0 + chain);
          q1 = (uint32_t)step;
          philox4x32_10(q0, q1, q2, q3, p.seed_lo, p.seed_hi);
          qthr = log1p(-((double)((((uint64_t)q2 << 32) | q3) >> 11) * 0x1.0p-53));
        }
        int src = t & 31;
        uint32_t r0 = __shfl_sync(FULL, q0, src), r1 = __shfl_sync(FULL, q1, src);
        thr = __shfl_sync(FULL, qthr, src);
        if (p.kind == QTX_LOCAL_FLIP) {
          pos = (int)__umulhi(r0, (uint32_t)N);
        } else {
          // k-th site holding the hopping particle
          uint32_t w = (lane < NW) ? myword : 0u;
          if (p.hop < 0) {
            w = ~w;
            int rem = N - lane * 32;
            if (lane >= NW) w = 0;
            else if (rem < 32) w &= (1u << rem) - 1u;
          }
          int cnt = __popc(w), incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
          }
          int nhop = __shfl_sync(FULL, incl, 31);
          int k = (int)__umulhi(r0, (uint32_t)nhop);
          uint32_t bal = __ballot_sync(FULL, incl > k);
          int L = __ffs(bal) - 1;
          int excl = __shfl_sync(FULL, incl - cnt, L);
          uint32_t wl = __shfl_sync(FULL, w, L);
          pos = L * 32 + (int)__fns(wl, 0, k - excl + 1);
          slot = (int)__umulhi(r1, (uint32_t)p.max_nb);
        }
      }
      int j0 = pos, j1 = -1;
      bool moved = true;
      if (p.kind == QTX_SPIN_EXCHANGE) {
        int nb = p.nbr[pos * p.max_nb + slot];
        if (nb < 0) nb = pos;
        j1 = nb;
        moved = spin_from_words(myword, j0) != spin_from_words(myword, j1);
      }
      bool acc = false;
      if (moved) {
        // new spin values at the flipped sites: s'_j = -s_j
        T s0 = (T)(-spin_from_words(myword, j0));
        const T* c0 = Wt + (size_t)j0 * M;
        T s1 = 0;
        const T* c1 = c0;
        if (j1 >= 0) {
          s1 = (T)(-spin_from_words(myword, j1));
          c1 = Wt + (size_t)j1 * M;
        }
        T ls = 0;
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
          int i = r * 32 + lane;
          if (i < M) {
            T d = c0[i] * s0 + c1[i] * s1;
            ls += lncosh(th[r] + T(2) * d);
          }
        }
        double la_new = (double)warp_sum(ls);
        acc = p.reweight * (la_new - la) > thr;  // |psi| == 0 cannot happen: log|cosh| >= 0
        if (acc) {
#pragma unroll
          for (int r = 0; r < RMAX; ++r) {
            int i = r * 32 + lane;
            if (i < M) {
              T d = c0[i] * s0 + c1[i] * s1;
              th[r] = th[r] + T(2) * d;
            }
          }
          la = la_new;
          if (lane == (j0 >> 5)) myword ^= 1u << (j0 & 31);
          if (j1 >= 0 && lane == (j1 >> 5)) myword ^= 1u << (j1 & 31);
          ++nacc;
        }
      }
      if (p.accept_log && lane == 0) p.accept_log[(size_t)t * p.ns + chain] = acc ? 1 : 0;
    }

    // write back the chain, its locally-updated amplitude and the direct-forward amplitude
    for (int w = 0; w < NW; ++w) {
      uint32_t word = __shfl_sync(FULL, myword, w);
      int j = w * 32 + lane;
      if (j < N) sp[j] = ((word >> lane) & 1u) ? 1 : -1;
    }
    if (lane == 0) {
      if (p.logabs_chain_out) p.logabs_chain_out[chain] = la;
      if (p.naccept_out) p.naccept_out[chain] = nacc;
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) th[r] = 0;
    for (int j = 0; j < N; ++j) {
      T sj = (T)spin_from_words(myword, j);
      const T* col = Wt + (size_t)j * M;
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        int i = r * 32 + lane;
        if (i < M) th[r] += col[i] * sj;
      }
    }
    lsum = 0;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      int i = r * 32 + lane;
      if (i < M) lsum += lncosh(th[r] + p.b[i]);
    }
    lsum = warp_sum(lsum);
    if (lane == 0 && p.logabs_out) p.logabs_out[chain] = (double)lsum;
  }
}

constexpr size_t kSmemBudget = 227 * 1024 - 1024;

template <typename T, int RMAX>
static int launch_sweep(const SweepParams<T>& p, cudaStream_t st) {
  const size_t wbytes = (size_t)p.N * p.M * sizeof(T);
  const bool wsmem = wbytes <= kSmemBudget;
  const int maxthreads = RMAX <= 8 ? 1024 : (sizeof(T) == 4 && RMAX <= 16 ? 896 : 512);
  const int sms = num_sms();
  int wpb = (int)((p.ns + sms - 1) / sms);
  if (wpb > maxthreads / 32) wpb = maxthreads / 32;
  if (wpb < 1) wpb = 1;
  int64_t nblk = (p.ns + wpb - 1) / wpb;
  int per_sm = wsmem ? 1 : (2048 / (wpb * 32) > 0 ? 2048 / (wpb * 32) : 1);
  if (nblk > (int64_t)sms * per_sm) nblk = (int64_t)sms * per_sm;
  if (wsmem) {
    auto k = rbm_sweep_kernel<T, RMAX, true>;
    QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes));
    k<<<(unsigned)nblk, wpb * 32, wbytes, st>>>(p);
  } else {
    rbm_sweep_kernel<T, RMAX, false><<<(unsigned)nblk, wpb * 32, 0, st>>>(p);
  }
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

// hidden-unit registers per lane: the smallest instantiated value >= ceil(M / 32) (unused registers still
// cost issue slots, so the grid of instantiations is fairly fine: M = 400 -> 13)
#define QTX_RBM_DISPATCH(LAUNCH, T, p, st)                         \
  do {                                                             \
    const int nreg_ = ((p).M + 31) / 32;                           \
    if (nreg_ <= 2) return LAUNCH<T, 2>(p, st);                    \
    if (nreg_ <= 4) return LAUNCH<T, 4>(p, st);                    \
    if (nreg_ <= 6) return LAUNCH<T, 6>(p, st);                    \
    if (nreg_ <= 8) return LAUNCH<T, 8>(p, st);                    \
    if (nreg_ <= 10) return LAUNCH<T, 10>(p, st);                  \
    if (nreg_ <= 13) return LAUNCH<T, 13>(p, st);                  \
    if (nreg_ <= 16) return LAUNCH<T, 16>(p, st);                  \
    if (nreg_ <= 20) return LAUNCH<T, 20>(p, st);                  \
    if (nreg_ <= 26) return LAUNCH<T, 26>(p, st);                  \
    if (nreg_ <= 32) return LAUNCH<T, 32>(p, st);                  \
  } while (0)

template <typename T>
static int sweep_dispatch(const SweepParams<T>& p, cudaStream_t st) {
  QTX_RBM_DISPATCH(launch_sweep, T, p, st);
  set_error("qtx_rbm_sweep: M=%d > 1024 hidden units is not supported", p.M);
  return QTX_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// fused local energy
// ---------------------------------------------------------------------------------------------
template <typename T>
struct OlocParams {
  const T* Wt;
  const T* b;
  int N, M;
  const int8_t* spins;
  int64_t ns;
  const double* coef;
  const uint16_t* sites;  // [nterms, 4]
  const uint8_t* ops;     // [nterms, 4]
  int nterms;
  double* eloc;
  int32_t* nconn;
  int terms_in_smem;
};

template <typename T, int RMAX, bool WSMEM>
__global__ void __launch_bounds__(RMAX <= 8 ? 1024 : (sizeof(T) == 4 && RMAX <= 16 ? 896 : 512), 1)
    rbm_oloc_kernel(OlocParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N, M = p.M;
  unsigned char* sm = smem_raw;
  const T* Wt = p.Wt;
  if (WSMEM) {
    T* Ws = reinterpret_cast<T*>(sm);
    stage_w(Ws, p.Wt, (size_t)N * M);
    Wt = Ws;
    sm += (((size_t)N * M * sizeof(T)) + 15) / 16 * 16;
  }
  const double* coef = p.coef;
  const uint2* sites = reinterpret_cast<const uint2*>(p.sites);
  const uint32_t* ops = reinterpret_cast<const uint32_t*>(p.ops);
  if (p.terms_in_smem) {
    double* c_s = reinterpret_cast<double*>(sm);
    uint2* s_s = reinterpret_cast<uint2*>(c_s + p.nterms);
    uint32_t* o_s = reinterpret_cast<uint32_t*>(s_s + p.nterms);
    for (int t = threadIdx.x; t < p.nterms; t += blockDim.x) {
      c_s[t] = p.coef[t];
      s_s[t] = sites[t];
      o_s[t] = ops[t];
    }
    coef = c_s; sites = s_s; ops = o_s;
    sm += (size_t)p.nterms * 20;
    sm = (unsigned char*)(((uintptr_t)sm + 15) / 16 * 16);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  int8_t* myspins = reinterpret_cast<int8_t*>(sm) + (size_t)warp * ((N + 15) / 16 * 16);
  __syncthreads();
  const int64_t nwarps = (int64_t)gridDim.x * wpb;

  for (int64_t s = (int64_t)blockIdx.x * wpb + warp; s < p.ns; s += nwarps) {
    const int8_t* sp = p.spins + s * N;
    __syncwarp();
    for (int j = lane; j < N; j += 32) myspins[j] = sp[j];
    __syncwarp();
    T th[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) th[r] = 0;
    for (int j = 0; j < N; ++j) {
      T sj = (T)myspins[j];
      const T* col = Wt + (size_t)j * M;
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        int i = r * 32 + lane;
        if (i < M) th[r] += col[i] * sj;
      }
    }
    T lsum = 0;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      int i = r * 32 + lane;
      if (i < M) {
        th[r] += p.b[i];
        lsum += lncosh(th[r]);
      }
    }
    const double la = (double)warp_sum(lsum);
    double e = 0.0, ediag = 0.0;
    int nconn = 0;
    // 32 terms are decoded at once (one per lane); the warp then visits only the valid off-diagonal ones
    for (int t0 = 0; t0 < p.nterms; t0 += 32) {
      const int t = t0 + lane;
      double c = 0.0;
      uint2 st = make_uint2(0u, 0u);
      uint32_t flipmask = 0;  // bit k set: site k of the term is flipped
      bool valid = false;
      if (t < p.nterms) {
        c = coef[t];
        st = sites[t];
        const uint32_t op4 = ops[t];
        const int site[4] = {(int)(st.x & 0xffff), (int)(st.x >> 16), (int)(st.y & 0xffff), (int)(st.y >> 16)};
        valid = true;
#pragma unroll
        for (int k = 3; k >= 0; --k) {  // right-most operator acts first (operator.py:107)
          const int op = (op4 >> (8 * k)) & 0xff;
          if (op == QTX_OP_NONE || op == QTX_OP_I) continue;
          const int sk = myspins[site[k]];
          if (op == QTX_OP_Z) {
            c = c * sk / 2;
          } else {
            if (op == QTX_OP_X) c = c / 2;
            else if (op == QTX_OP_P) valid = valid && (sk < 0);
            else valid = valid && (sk > 0);
            flipmask |= 1u << k;
          }
        }
        if (flipmask == 0) {
          ediag += c;
          valid = false;
        } else {
          valid = valid && fabs(c) > 1e-8;  // NaN or isclose(H, 0) (operator.py:154)
        }
      }
      uint32_t todo = __ballot_sync(FULL, valid);
      nconn += __popc(todo);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const double cc = __shfl_sync(FULL, c, src);
        const uint32_t sx = __shfl_sync(FULL, st.x, src), sy = __shfl_sync(FULL, st.y, src);
        const uint32_t fm = __shfl_sync(FULL, flipmask, src);
        const int site[4] = {(int)(sx & 0xffff), (int)(sx >> 16), (int)(sy & 0xffff), (int)(sy >> 16)};
        const T* cp[4];
        T sg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool f = (fm >> k) & 1u;
          cp[k] = f ? Wt + (size_t)site[k] * M : nullptr;
          sg[k] = f ? (T)(-myspins[site[k]]) : T(0);  // new spin value at the flipped site
        }
        T ls = 0;
        if (fm == 3u) {  // two-site exchange / hop terms (Heisenberg, J1-J2)
#pragma unroll
          for (int r = 0; r < RMAX; ++r) {
            const int i = r * 32 + lane;
            if (i < M) ls += lncosh(th[r] + T(2) * (cp[0][i] * sg[0] + cp[1][i] * sg[1]));
          }
        } else if (fm == 1u) {  // single-site flips (transverse field)
#pragma unroll
          for (int r = 0; r < RMAX; ++r) {
            const int i = r * 32 + lane;
            if (i < M) ls += lncosh(th[r] + T(2) * (cp[0][i] * sg[0]));
          }
        } else {
#pragma unroll
          for (int r = 0; r < RMAX; ++r) {
            const int i = r * 32 + lane;
            if (i < M) {
              T d = 0;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (cp[k]) d += cp[k][i] * sg[k];
              ls += lncosh(th[r] + T(2) * d);
            }
          }
        }
        const double la_new = (double)warp_sum(ls);
        e += cc * exp(la_new - la);
      }
    }
    e += warp_sum(ediag);
    if (lane == 0) {
      p.eloc[s] = e;
      if (p.nconn) p.nconn[s] = nconn;
    }
  }
}

template <typename T, int RMAX>
static int launch_oloc(OlocParams<T> p, cudaStream_t st) {
  const size_t wbytes = (((size_t)p.N * p.M * sizeof(T)) + 15) / 16 * 16;
  const size_t tbytes = ((size_t)p.nterms * 20 + 15) / 16 * 16;
  const int maxthreads = RMAX <= 8 ? 1024 : (sizeof(T) == 4 && RMAX <= 16 ? 896 : 512);
  const int sms = num_sms();
  int wpb = (int)((p.ns + sms - 1) / sms);
  if (wpb > maxthreads / 32) wpb = maxthreads / 32;
  if (wpb < 1) wpb = 1;
  const size_t sbytes = (size_t)wpb * ((p.N + 15) / 16 * 16);
  const bool wsmem = wbytes + sbytes + 64 <= kSmemBudget;
  size_t smem = (wsmem ? wbytes : 0) + sbytes + 64;
  p.terms_in_smem = (smem + tbytes <= kSmemBudget) ? 1 : 0;
  if (p.terms_in_smem) smem += tbytes;
  int64_t nblk = (p.ns + wpb - 1) / wpb;
  int per_sm = wsmem ? 1 : (2048 / (wpb * 32) > 0 ? 2048 / (wpb * 32) : 1);
  if (nblk > (int64_t)sms * per_sm) nblk = (int64_t)sms * per_sm;
  if (wsmem) {
    auto k = rbm_oloc_kernel<T, RMAX, true>;
    QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)nblk, wpb * 32, smem, st>>>(p);
  } else {
    auto k = rbm_oloc_kernel<T, RMAX, false>;
    QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)nblk, wpb * 32, smem, st>>>(p);
  }
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

template <typename T>
static int oloc_dispatch(const OlocParams<T>& p, cudaStream_t st) {
  QTX_RBM_DISPATCH(launch_oloc, T, p, st);
  set_error("qtx_rbm_oloc: M=%d > 1024 hidden units is not supported", p.M);
  return QTX_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// ref_forward: psi of connected configurations from the parent's theta
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rbm_ref_forward_kernel(const T* __restrict__ W, int N, int M, const T* __restrict__ theta,
                                       const int8_t* __restrict__ s_old, const int8_t* __restrict__ s_new,
                                       const int32_t* __restrict__ segment, int64_t nconn, int nflips,
                                       double* __restrict__ logabs_out) {
  const int lane = threadIdx.x & 31;
  int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= nconn) return;
  int seg = segment[c];
  if (seg < 0) {
    // padding entry: the reference gathers row -1 (jnp negative index); its value is multiplied by
    // H = 0 downstream.  We emit logabs = 0.
    if (lane == 0) logabs_out[c] = 0.0;
    return;
  }
  const int8_t* so = s_old + (size_t)seg * N;
  const int8_t* sn = s_new + c * N;
  // first `nflips` differing sites in ascending order, missing ones padded with site 0
  int idx[QTX_MAX_TERM_SITES] = {0, 0, 0, 0};
  int found = 0;
  for (int j0 = 0; j0 < N && found < nflips; j0 += 32) {
    int j = j0 + lane;
    uint32_t bal = __ballot_sync(FULL, j < N && so[j] != sn[j]);
    while (bal && found < nflips) {
      int bpos = __ffs(bal) - 1;
      idx[found++] = j0 + bpos;
      bal &= bal - 1;
    }
  }
  const T* th = theta + (size_t)seg * M;
  T ls = 0;
  for (int i = lane; i < M; i += 32) {
    T d = 0;
    for (int k = 0; k < nflips; ++k) d += W[(size_t)i * N + idx[k]] * (T)sn[idx[k]];
    ls += lncosh(th[i] + T(2) * d);
  }
  ls = warp_sum(ls);
  if (lane == 0) logabs_out[c] = (double)ls;
}

// ---------------------------------------------------------------------------------------------
// Jacobian: O[s, i*N+j] = tanh(theta_i) s_j ; O[s, M*N+i] = tanh(theta_i)
// one CTA per (sample, column tile); fully coalesced 16-byte stores
// ---------------------------------------------------------------------------------------------
template <typename OutT> struct OutVec2;
template <> struct OutVec2<double> { using type = double2; };
template <> struct OutVec2<float> { using type = float2; };

template <typename T, typename OutT>
__global__ void __launch_bounds__(256) rbm_jacobian_kernel(const T* __restrict__ W, const T* __restrict__ b, int N,
                                                           int M, const int8_t* __restrict__ spins, int64_t ns,
                                                           OutT* __restrict__ out, int64_t ld,
                                                           const double* __restrict__ col_mean,
                                                           const double* __restrict__ row_scale,
                                                           const T* __restrict__ tanh_table) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* t_s = reinterpret_cast<T*>(smem_raw);          // [M] tanh(theta)
  T* s_s = t_s + ((M + 3) / 4 * 4);                  // [N] spins as T
  const int64_t s = blockIdx.x;
  const int8_t* sp = spins + s * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) s_s[j] = (T)sp[j];
  __syncthreads();
  if (tanh_table) {
    for (int i = threadIdx.x; i < M; i += blockDim.x) t_s[i] = tanh_table[s * M + i];
  } else {
    // same summation order as rbm_tanh_kernel (sites ascending, bias last), so the rows written here are
    // bitwise consistent with the column means of qtx_rbm_jacobian_colmean
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      const T* wr = W + (size_t)i * N;
      T acc = 0;
      for (int j = 0; j < N; ++j) acc += wr[j] * s_s[j];
      t_s[i] = tanh(acc + b[i]);
    }
  }
  __syncthreads();
  const double scale = row_scale ? row_scale[s] : 1.0;
  const int64_t MN = (int64_t)M * N, NP = MN + M;
  OutT* row = out + s * ld;
  using V2 = typename OutVec2<OutT>::type;
  const int P2 = N >> 1;  // column pairs per hidden unit
  const bool fast = ((N & 1) == 0) && P2 <= (int)blockDim.x && ((reinterpret_cast<uintptr_t>(row) & (sizeof(V2) - 1)) == 0) &&
                    (!col_mean || (reinterpret_cast<uintptr_t>(col_mean) & 15) == 0);
  if (fast) {
    // thread <-> fixed column pair (2 jp, 2 jp + 1); G hidden units are written per block iteration as one
    // contiguous run of G * N elements: fully coalesced 16-byte stores, ~10 instructions per pair.
    const int G = blockDim.x / P2;
    const int g = threadIdx.x / P2, jp = threadIdx.x - g * P2;
    if (g < G) {
      const T s0 = s_s[2 * jp], s1 = s_s[2 * jp + 1];
      for (int i = g; i < M; i += G) {
        const T t = t_s[i];
        double v0 = (double)(t * s0), v1 = (double)(t * s1);
        const int64_t k = (int64_t)i * N + 2 * jp;
        if (col_mean) {
          const double2 m = *reinterpret_cast<const double2*>(col_mean + k);
          v0 -= m.x;
          v1 -= m.y;
        }
        V2 o;
        o.x = (OutT)(v0 * scale);
        o.y = (OutT)(v1 * scale);
        *reinterpret_cast<V2*>(row + k) = o;
      }
    }
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      double v = (double)t_s[i];
      if (col_mean) v -= col_mean[MN + i];
      row[MN + i] = (OutT)(v * scale);
    }
  } else {
    for (int64_t k = threadIdx.x; k < NP; k += blockDim.x) {
      double val = (k < MN) ? (double)(t_s[k / N] * s_s[k % N]) : (double)t_s[k - MN];
      if (col_mean) val -= col_mean[k];
      row[k] = (OutT)(val * scale);
    }
  }
}

// tanh(theta) table [ns, M] (model dtype) for the column-mean kernel: warp per sample, W^T [N, M] read
// through L1/L2 with coalesced rows (hidden unit i = r*32 + lane), theta in registers.
template <typename T, int RMAX>
__global__ void __launch_bounds__(256) rbm_tanh_kernel(const T* __restrict__ Wt, const T* __restrict__ b, int N, int M,
                                                       const int8_t* __restrict__ spins, int64_t ns,
                                                       T* __restrict__ t_out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= ns) return;
  const int8_t* sp = spins + s * N;
  T th[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; ++r) th[r] = 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int jj = j0 + lane;
    const int mine = jj < N ? sp[jj] : 0;
    const int jn = min(32, N - j0);
    for (int q = 0; q < jn; ++q) {
      const T sj = (T)__shfl_sync(FULL, mine, q);
      const T* col = Wt + (size_t)(j0 + q) * M;
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        const int i = r * 32 + lane;
        if (i < M) th[r] += col[i] * sj;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RMAX; ++r) {
    const int i = r * 32 + lane;
    if (i < M) t_out[s * M + i] = tanh(th[r] + b[i]);
  }
}

// mean[i*N+j] = (1/ns) sum_s w_s t[s,i] spins[s,j];  mean[M*N+i] = (1/ns) sum_s w_s t[s,i]
// CTA = 16 hidden units x all (N+1) columns x one chunk of 256 samples staged in shared memory;
// float64 accumulation of the exact products, chunks combined with atomicAdd(double).
constexpr int kCmI = 16;
template <typename T>
__global__ void __launch_bounds__(256) rbm_colmean_kernel(const T* __restrict__ t, const int8_t* __restrict__ spins,
                                                          int N, int M, int64_t ns, const double* __restrict__ weight,
                                                          double* __restrict__ mean_out, int kCmS) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* t_s = reinterpret_cast<T*>(smem_raw);                        // [kCmS][kCmI]
  T* s_s = t_s + kCmS * kCmI;                                     // [kCmS][N+1] (last column = 1)
  double* w_s = reinterpret_cast<double*>(s_s + (size_t)kCmS * (N + 1) + ((kCmS * (N + 1)) & 1));
  const int i0 = blockIdx.x * kCmI;
  const int64_t s0 = (int64_t)blockIdx.y * kCmS;
  const int nsl = (int)min((int64_t)kCmS, ns - s0);
  const int NC = N + 1;
  for (int e = threadIdx.x; e < kCmS * kCmI; e += blockDim.x) {
    int sl = e / kCmI, il = e % kCmI;
    t_s[e] = (sl < nsl && i0 + il < M) ? t[(s0 + sl) * M + i0 + il] : T(0);
  }
  for (int e = threadIdx.x; e < kCmS * NC; e += blockDim.x) {
    int sl = e / NC, j = e % NC;
    s_s[e] = (sl < nsl) ? (j < N ? (T)spins[(s0 + sl) * N + j] : T(1)) : T(0);
  }
  for (int e = threadIdx.x; e < kCmS; e += blockDim.x) w_s[e] = (e < nsl) ? (weight ? weight[s0 + e] : 1.0) : 0.0;
  __syncthreads();
  const int nout = kCmI * NC;
  const double inv = 1.0 / (double)ns;
  for (int o = threadIdx.x; o < nout; o += blockDim.x) {
    const int il = o / NC, j = o % NC;
    if (i0 + il >= M) continue;
    double acc = 0.0;
#pragma unroll 4
    for (int sl = 0; sl < kCmS; ++sl) acc += w_s[sl] * (double)(t_s[sl * kCmI + il] * s_s[sl * NC + j]);
    const int64_t k = (j < N) ? (int64_t)(i0 + il) * N + j : (int64_t)M * N + i0 + il;
    atomicAdd(mean_out + k, acc * inv);
  }
}

}  // namespace qtx

// ===============================================================================================
// C ABI
// ===============================================================================================
using namespace qtx;

extern "C" size_t qtx_rbm_workspace_size(int model_dtype, int N, int M) {
  size_t es = model_dtype == QTX_F64 ? 8 : 4;
  return ((size_t)N * M * es + 255) / 256 * 256;
}

extern "C" int qtx_rbm_forward(int model_dtype, const void* W, const void* b, int N, int M, const int8_t* spins,
                               int64_t ns, void* theta_out, double* logabs_out, qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(W && b && spins && N > 0 && M > 0 && ns >= 0, QTX_ERR_INVALID, "qtx_rbm_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)((ns + 7) / 8);
  if (model_dtype == QTX_F32)
    rbm_forward_kernel<float><<<grid, 256, 0, st>>>((const float*)W, (const float*)b, N, M, spins, ns,
                                                    (float*)theta_out, logabs_out);
  else if (model_dtype == QTX_F64)
    rbm_forward_kernel<double><<<grid, 256, 0, st>>>((const double*)W, (const double*)b, N, M, spins, ns,
                                                     (double*)theta_out, logabs_out);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_forward: bad dtype %d", model_dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

template <typename T>
static int rbm_sweep_impl(const void* W, const void* b, int N, int M, int8_t* spins, int64_t ns, int nsweeps, int kind,
                          const int32_t* nbr, int max_nb, int hop, double reweight, const int32_t* inj_pos,
                          const int32_t* inj_slot, const double* inj_u, uint64_t seed, uint64_t step0, uint64_t chain0,
                          double* logabs_out, double* logabs_chain_out, int32_t* naccept_out, uint8_t* accept_log,
                          void* ws, cudaStream_t st) {
  int rc = transpose_w<T>(W, ws, N, M, st);
  if (rc) return rc;
  SweepParams<T> p;
  p.Wt = (const T*)ws; p.b = (const T*)b; p.N = N; p.M = M; p.spins = spins; p.ns = ns;
  p.nsweeps = nsweeps; p.kind = kind; p.nbr = nbr; p.max_nb = max_nb; p.hop = hop; p.reweight = reweight;
  p.inj_pos = inj_pos; p.inj_slot = inj_slot; p.inj_u = inj_u;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.step0 = step0; p.chain0 = chain0;
  p.logabs_out = logabs_out; p.logabs_chain_out = logabs_chain_out; p.naccept_out = naccept_out;
  p.accept_log = accept_log;
  return sweep_dispatch<T>(p, st);
}

extern "C" int qtx_rbm_sweep(int model_dtype, const void* W, const void* b, int N, int M, int8_t* spins, int64_t ns,
                             int nsweeps, int kind, const int32_t* nbr_table, int max_nb, int hop, double reweight,
                             const int32_t* inj_pos, const int32_t* inj_slot, const double* inj_u, uint64_t seed,
                             uint64_t step0, uint64_t chain0, double* logabs_out, double* logabs_chain_out,
                             int32_t* naccept_out, uint8_t* accept_log, void* workspace, size_t workspace_bytes,
                             qtx_stream_t stream) {
  QTX_REQUIRE(W && b && spins && N > 0 && M > 0 && ns >= 0 && nsweeps >= 0, QTX_ERR_INVALID,
              "qtx_rbm_sweep: bad argument");
  QTX_REQUIRE(N <= 1024, QTX_ERR_UNSUPPORTED, "qtx_rbm_sweep: N=%d > 1024 sites is not supported", N);
  QTX_REQUIRE(kind == QTX_LOCAL_FLIP || kind == QTX_SPIN_EXCHANGE, QTX_ERR_INVALID, "qtx_rbm_sweep: bad kind %d", kind);
  if (kind == QTX_SPIN_EXCHANGE)
    QTX_REQUIRE(nbr_table && max_nb > 0 && (hop == 1 || hop == -1), QTX_ERR_INVALID,
                "qtx_rbm_sweep: exchange needs a neighbour table and hop = +-1");
  if (inj_u) QTX_REQUIRE(inj_pos && (kind == QTX_LOCAL_FLIP || inj_slot), QTX_ERR_INVALID,
                         "qtx_rbm_sweep: injected randoms need inj_pos (and inj_slot for exchange)");
  QTX_REQUIRE(workspace && workspace_bytes >= qtx_rbm_workspace_size(model_dtype, N, M), QTX_ERR_INVALID,
              "qtx_rbm_sweep: workspace too small");
  if (ns == 0) return QTX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32)
    return rbm_sweep_impl<float>(W, b, N, M, spins, ns, nsweeps, kind, nbr_table, max_nb, hop, reweight, inj_pos,
                                 inj_slot, inj_u, seed, step0, chain0, logabs_out, logabs_chain_out, naccept_out,
                                 accept_log, workspace, st);
  if (model_dtype == QTX_F64)
    return rbm_sweep_impl<double>(W, b, N, M, spins, ns, nsweeps, kind, nbr_table, max_nb, hop, reweight, inj_pos,
                                  inj_slot, inj_u, seed, step0, chain0, logabs_out, logabs_chain_out, naccept_out,
                                  accept_log, workspace, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_sweep: bad dtype %d", model_dtype);
}

template <typename T>
static int rbm_oloc_impl(const void* W, const void* b, int N, int M, const int8_t* spins, int64_t ns,
                         const double* coef, const uint16_t* sites, const uint8_t* ops, int nterms, double* eloc,
                         int32_t* nconn, void* ws, cudaStream_t st) {
  int rc = transpose_w<T>(W, ws, N, M, st);
  if (rc) return rc;
  OlocParams<T> p;
  p.Wt = (const T*)ws; p.b = (const T*)b; p.N = N; p.M = M; p.spins = spins; p.ns = ns;
  p.coef = coef; p.sites = sites; p.ops = ops; p.nterms = nterms; p.eloc = eloc; p.nconn = nconn;
  p.terms_in_smem = 0;
  return oloc_dispatch<T>(p, st);
}

extern "C" int qtx_rbm_oloc(int model_dtype, const void* W, const void* b, int N, int M, const int8_t* spins,
                            int64_t ns, const double* term_coef, const uint16_t* term_sites, const uint8_t* term_ops,
                            int nterms, double* eloc_out, int32_t* nconn_out, void* workspace, size_t workspace_bytes,
                            qtx_stream_t stream) {
  if (ns == 0) return QTX_OK;
  QTX_REQUIRE(W && b && spins && eloc_out && N > 0 && M > 0 && ns >= 0 && nterms >= 0, QTX_ERR_INVALID,
              "qtx_rbm_oloc: bad argument");
  QTX_REQUIRE(nterms == 0 || (term_coef && term_sites && term_ops), QTX_ERR_INVALID, "qtx_rbm_oloc: null term table");
  QTX_REQUIRE(workspace && workspace_bytes >= qtx_rbm_workspace_size(model_dtype, N, M), QTX_ERR_INVALID,
              "qtx_rbm_oloc: workspace too small");
  if (ns == 0) return QTX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32)
    return rbm_oloc_impl<float>(W, b, N, M, spins, ns, term_coef, term_sites, term_ops, nterms, eloc_out, nconn_out,
                                workspace, st);
  if (model_dtype == QTX_F64)
    return rbm_oloc_impl<double>(W, b, N, M, spins, ns, term_coef, term_sites, term_ops, nterms, eloc_out, nconn_out,
                                 workspace, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_oloc: bad dtype %d", model_dtype);
}

extern "C" int qtx_rbm_ref_forward(int model_dtype, const void* W, int N, int M, const void* theta,
                                   const int8_t* s_old, int64_t ns, const int8_t* s_new, const int32_t* segment,
                                   int64_t nconn, int nflips, double* logabs_out, qtx_stream_t stream) {
  QTX_REQUIRE(W && theta && s_old && s_new && segment && logabs_out && N > 0 && M > 0, QTX_ERR_INVALID,
              "qtx_rbm_ref_forward: bad argument");
  QTX_REQUIRE(nflips >= 1 && nflips <= QTX_MAX_TERM_SITES, QTX_ERR_UNSUPPORTED,
              "qtx_rbm_ref_forward: nflips=%d outside 1..%d", nflips, QTX_MAX_TERM_SITES);
  (void)ns;
  if (nconn == 0) return QTX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)((nconn + 7) / 8);
  if (model_dtype == QTX_F32)
    rbm_ref_forward_kernel<float><<<grid, 256, 0, st>>>((const float*)W, N, M, (const float*)theta, s_old, s_new,
                                                        segment, nconn, nflips, logabs_out);
  else if (model_dtype == QTX_F64)
    rbm_ref_forward_kernel<double><<<grid, 256, 0, st>>>((const double*)W, N, M, (const double*)theta, s_old, s_new,
                                                         segment, nconn, nflips, logabs_out);
  else QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_ref_forward: bad dtype %d", model_dtype);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

template <typename T, typename OutT>
static int jac_launch(const void* W, const void* b, int N, int M, const int8_t* spins, int64_t ns, void* out,
                      int64_t ld, const double* mean, const double* scale, const void* table, cudaStream_t st) {
  size_t smem = ((size_t)(M + 3) / 4 * 4 + N) * sizeof(T) + 16;
  QTX_REQUIRE(smem <= kSmemBudget, QTX_ERR_UNSUPPORTED, "qtx_rbm_jacobian: M+N too large for shared memory");
  auto k = rbm_jacobian_kernel<T, OutT>;
  if (smem > 48 * 1024) QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(unsigned)ns, 256, smem, st>>>((const T*)W, (const T*)b, N, M, spins, ns, (OutT*)out, ld, mean, scale,
                                     (const T*)table);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_rbm_jacobian(int model_dtype, const void* W, const void* b, int N, int M, const int8_t* spins,
                                int64_t ns, int out_dtype, void* out, int64_t ld, const double* col_mean,
                                const double* row_scale, const void* tanh_table, qtx_stream_t stream) {
  QTX_REQUIRE(W && b && spins && out && N > 0 && M > 0 && ns >= 0, QTX_ERR_INVALID, "qtx_rbm_jacobian: bad argument");
  QTX_REQUIRE(ld >= (int64_t)M * N + M, QTX_ERR_INVALID, "qtx_rbm_jacobian: ld smaller than the parameter count");
  QTX_REQUIRE(ns < (1ll << 31), QTX_ERR_UNSUPPORTED, "qtx_rbm_jacobian: ns too large");
  if (ns == 0) return QTX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32 && out_dtype == QTX_F64)
    return jac_launch<float, double>(W, b, N, M, spins, ns, out, ld, col_mean, row_scale, tanh_table, st);
  if (model_dtype == QTX_F32 && out_dtype == QTX_F32)
    return jac_launch<float, float>(W, b, N, M, spins, ns, out, ld, col_mean, row_scale, tanh_table, st);
  if (model_dtype == QTX_F64 && out_dtype == QTX_F64)
    return jac_launch<double, double>(W, b, N, M, spins, ns, out, ld, col_mean, row_scale, tanh_table, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_jacobian: unsupported dtype pair (%d -> %d)", model_dtype, out_dtype);
}

extern "C" size_t qtx_rbm_colmean_workspace_size(int model_dtype, int N, int M, int64_t ns) {
  size_t es = model_dtype == QTX_F64 ? 8 : 4;
  return ((size_t)ns * M * es + 255) / 256 * 256 + ((size_t)N * M * es + 255) / 256 * 256;
}

template <typename T, int RMAX>
static int tanh_launch(const T* Wt, const T* b, int N, int M, const int8_t* spins, int64_t ns, T* t, cudaStream_t st) {
  rbm_tanh_kernel<T, RMAX><<<(unsigned)((ns + 7) / 8), 256, 0, st>>>(Wt, b, N, M, spins, ns, t);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

template <typename T>
static int colmean_impl(const void* W, const void* b, int N, int M, const int8_t* spins, int64_t ns,
                        const double* weight, double* mean_out, void* tanh_out, void* ws, cudaStream_t st) {
  T* t = tanh_out ? (T*)tanh_out : (T*)ws;
  T* Wt = (T*)((char*)ws + ((size_t)ns * M * sizeof(T) + 255) / 256 * 256);
  int rc = transpose_w<T>(W, Wt, N, M, st);
  if (rc) return rc;
  if (M <= 64) rc = tanh_launch<T, 2>(Wt, (const T*)b, N, M, spins, ns, t, st);
  else if (M <= 128) rc = tanh_launch<T, 4>(Wt, (const T*)b, N, M, spins, ns, t, st);
  else if (M <= 256) rc = tanh_launch<T, 8>(Wt, (const T*)b, N, M, spins, ns, t, st);
  else if (M <= 512) rc = tanh_launch<T, 16>(Wt, (const T*)b, N, M, spins, ns, t, st);
  else if (M <= 1024) rc = tanh_launch<T, 32>(Wt, (const T*)b, N, M, spins, ns, t, st);
  else QTX_REQUIRE(false, QTX_ERR_UNSUPPORTED, "qtx_rbm_jacobian_colmean: M=%d > 1024 is not supported", M);
  if (rc) return rc;
  const int64_t NP = (int64_t)M * N + M;
  QTX_CUDA(cudaMemsetAsync(mean_out, 0, NP * sizeof(double), st));
  int kCmS = 256;  // samples staged per CTA; shrink (even values) until the tile fits in shared memory
  auto smem_of = [&](int sc) { return ((size_t)sc * kCmI + (size_t)sc * (N + 1) + 1) * sizeof(T) + sc * sizeof(double) + 16; };
  while (kCmS > 16 && smem_of(kCmS) > 100 * 1024) kCmS -= 16;
  const size_t smem = smem_of(kCmS);
  QTX_REQUIRE(smem <= kSmemBudget, QTX_ERR_UNSUPPORTED, "qtx_rbm_jacobian_colmean: N too large for shared memory");
  auto k = rbm_colmean_kernel<T>;
  if (smem > 48 * 1024) QTX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((M + kCmI - 1) / kCmI), (unsigned)((ns + kCmS - 1) / kCmS));
  k<<<grid, 256, smem, st>>>(t, spins, N, M, ns, weight, mean_out, kCmS);
  QTX_LAUNCH_CHECK();
  return QTX_OK;
}

extern "C" int qtx_rbm_jacobian_colmean(int model_dtype, const void* W, const void* b, int N, int M,
                                        const int8_t* spins, int64_t ns, const double* weight, double* mean_out,
                                        void* tanh_out, void* workspace, size_t workspace_bytes, qtx_stream_t stream) {
  QTX_REQUIRE(W && b && spins && mean_out && N > 0 && M > 0 && ns > 0, QTX_ERR_INVALID,
              "qtx_rbm_jacobian_colmean: bad argument");
  QTX_REQUIRE(workspace && workspace_bytes >= qtx_rbm_colmean_workspace_size(model_dtype, N, M, ns), QTX_ERR_INVALID,
              "qtx_rbm_jacobian_colmean: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (model_dtype == QTX_F32) return colmean_impl<float>(W, b, N, M, spins, ns, weight, mean_out, tanh_out, workspace, st);
  if (model_dtype == QTX_F64) return colmean_impl<double>(W, b, N, M, spins, ns, weight, mean_out, tanh_out, workspace, st);
  QTX_REQUIRE(false, QTX_ERR_INVALID, "qtx_rbm_jacobian_colmean: bad dtype %d", model_dtype);
}
