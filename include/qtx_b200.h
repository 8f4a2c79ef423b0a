/*
 * qtx_b200.h -- C ABI of libqtx_b200.so: the B200 (sm_100a) kernels behind the quantax
 * VMC hot path (Metropolis sweep -> Operator.Oloc -> Variational.jacobian -> SR/MinSR solve).
 *
 * The reference (ChenAo-Phys/quantax v0.2.1) is pure JAX and has no FFI boundary on this
 * path; each entry point below replaces the body of one reference Python function (cited as
 * file:line under /root/reference) and is what an XLA-FFI / ctypes binding binds
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - every function only enqueues work on `stream` (a cudaStream_t); no implicit
 *     synchronisation and no allocation: scratch memory is passed in by the caller after a
 *     *_workspace_size query.  Exceptions are documented per function (eigh handle setup).
 *   - all pointers are DEVICE pointers unless the name ends in `_host`; arrays are dense
 *     row-major.
 *   - return value: 0 on success, negative qtx_status otherwise; qtx_last_error() returns a
 *     thread-local message.
 *   - dtype arguments take QTX_F32 / QTX_F64.  `model_dtype` is the parameter / internal
 *     arithmetic type of the network (reference default float32), amplitudes psi travel as
 *     float64 (mult, expo) pairs with psi = mult * exp(expo): LogArray(sign, logabs) for
 *     RBM_Dense, ScaleArray(significand, exponent) for ResConv
 *     (quantax/utils/big_array.py:154,407; cast at quantax/state/variational.py:266).
 *   - spins are int8 +-1, [ns, N] (quantax/sampler/samples.py:49).
 */
#ifndef QTX_B200_H
#define QTX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* qtx_stream_t; /* cudaStream_t */

enum qtx_status {
  QTX_OK = 0,
  QTX_ERR_INVALID = -1,     /* bad argument */
  QTX_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  QTX_ERR_UNSUPPORTED = -3, /* size or mode outside the implemented range */
  QTX_ERR_SOLVER = -4       /* cuSOLVER failure or non-convergence */
};

enum qtx_dtype { QTX_F32 = 0, QTX_F64 = 1, QTX_I32 = 2 /* collectives only */ };
enum qtx_reduce_op { QTX_SUM = 0, QTX_MAX = 1 };
typedef void* qtx_comm_t; /* communicator of qtx_comm_init / qtx_comm_adopt */

/* proposal kinds of the Metropolis sweep */
enum qtx_proposal {
  QTX_LOCAL_FLIP = 0,   /* quantax/sampler/common_samplers.py:14-33  */
  QTX_SPIN_EXCHANGE = 1 /* quantax/sampler/common_samplers.py:58-162 (a.k.a. NeighborExchange) */
};

const char* qtx_last_error(void);
int qtx_abi_version(void);
/* number of kernels launched by this library on the calling thread since the last reset */
int64_t qtx_launch_count(void);
void qtx_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------
 * Hamiltonian term table (host -> device), replaces Operator.jax_op_list
 * (quantax/operator/operator.py:219-236).  A term is a coefficient and up to 4 (site, op)
 * pairs applied right-to-left (operator.py:107); op codes: 0 none, 1 'z', 2 'x', 3 '+', 4 '-',
 * 5 'I'.  Terms are stored in op-list order; `nflips[t]` is the number of x/+/- factors.
 * ------------------------------------------------------------------------------------------ */
#define QTX_MAX_TERM_SITES 4
#define QTX_MAX_PEERS 8 /* ranks of one NVSwitch node in the fused Gram + exchange path */
enum qtx_opcode { QTX_OP_NONE = 0, QTX_OP_Z = 1, QTX_OP_X = 2, QTX_OP_P = 3, QTX_OP_M = 4, QTX_OP_I = 5 };

/* ------------------------------------------------------------------------------------------
 * RBM_Dense: psi(s) = prod_i cosh(theta_i), theta = W s + b
 *   W [M, N] row-major (eqx Linear.weight), b [M]; flat parameter order [W.ravel(), b]
 *   (quantax/model/shallow_nets.py:35-126).
 * ------------------------------------------------------------------------------------------ */

/* theta = W s + b and logabs = sum_i log|cosh theta_i| (sign is always +1 for real parameters).
 * Replaces SingleDense.init_internal (shallow_nets.py:81-85) and the direct forward
 * Variational.__call__ (quantax/state/variational.py:325-347) for RBM_Dense.
 * theta_out [ns, M] in model dtype (nullable); logabs_out float64 [ns] (nullable). */
int qtx_rbm_forward(int model_dtype, const void* W, const void* b, int N, int M,
                    const int8_t* spins, int64_t ns, void* theta_out, double* logabs_out,
                    qtx_stream_t stream);

/* Scratch bytes needed by qtx_rbm_sweep / qtx_rbm_oloc (transposed copy of W). */
size_t qtx_rbm_workspace_size(int model_dtype, int N, int M);

/* Whole Metropolis sweep (all `nsweeps` steps, all chains) in ONE launch.
 * Replaces Metropolis._partial_sweep / _single_sweep / _update
 * (quantax/sampler/metropolis.py:246-322) with RefModel local updates
 * (shallow_nets.py:87-108) for LocalFlip / SpinExchange proposals.
 *
 *   spins       [ns, N] in/out: chain state
 *   nbr_table   [N, max_nb] int32, -1 padded (common_samplers.py:36-55); exchange only
 *   hop         +1 / -1 : the hopping particle (common_samplers.py:138-142); exchange only
 *   reweight    exponent n of the sampled |psi|^n (metropolis.py:300)
 *   Randoms: if inj_u != NULL the proposals are INJECTED (parity mode): inj_pos [nsweeps, ns]
 *   is the flipped site (LocalFlip) or the chosen particle site (exchange), inj_slot
 *   [nsweeps, ns] the neighbour-table column (exchange), inj_u [nsweeps, ns] the uniforms of
 *   metropolis.py:303.  Otherwise Philox4x32-10 with key=seed, counter=(chain0+chain,
 *   step0+t) is used in-kernel.
 *   Outputs: logabs_out [ns] = direct-forward amplitude of the final chains
 *   (metropolis.py:201-213), logabs_chain_out [ns] (nullable) = the locally updated value
 *   (for the drift check), naccept_out int32 [ns] (nullable), accept_log uint8
 *   [nsweeps, ns] (nullable). */
int qtx_rbm_sweep(int model_dtype, const void* W, const void* b, int N, int M,
                  int8_t* spins, int64_t ns, int nsweeps, int kind,
                  const int32_t* nbr_table, int max_nb, int hop, double reweight,
                  const int32_t* inj_pos, const int32_t* inj_slot, const double* inj_u,
                  uint64_t seed, uint64_t step0, uint64_t chain0,
                  double* logabs_out, double* logabs_chain_out, int32_t* naccept_out,
                  uint8_t* accept_log, void* workspace, size_t workspace_bytes,
                  qtx_stream_t stream);

/* Local energies for a spin Hamiltonian, fused: enumerate connected configurations, evaluate
 * psi(s')/psi(s) by local updates from theta(s), reduce.  Replaces Operator.Oloc
 * (quantax/operator/operator.py:510-562) incl. _apply_diag/_apply_off_diag/_get_conn/
 * _get_Olocx (operator.py:81-184) and Variational.ref_forward (variational.py:301-323).
 * Term table (device): coef float64 [nterms], sites uint16 [nterms, 4], ops uint8 [nterms, 4].
 * eloc_out float64 [ns]; nconn_out int32 [ns] (nullable) = number of valid connections. */
int qtx_rbm_oloc(int model_dtype, const void* W, const void* b, int N, int M,
                 const int8_t* spins, int64_t ns,
                 const double* term_coef, const uint16_t* term_sites, const uint8_t* term_ops,
                 int nterms, double* eloc_out, int32_t* nconn_out,
                 void* workspace, size_t workspace_bytes, qtx_stream_t stream);

/* psi of connected configurations given their parent sample (Variational.ref_forward,
 * variational.py:387-422 / 301-311): theta [ns, M] model dtype, s_old [ns, N],
 * s_new [nconn, N], segment int32 [nconn] (-1 = padding -> output 0), nflips as in the
 * reference (indices found by comparing s_new with s_old[segment], missing ones padded with
 * site 0, shallow_nets.py:101).  logabs_out float64 [nconn]. */
int qtx_rbm_ref_forward(int model_dtype, const void* W, int N, int M, const void* theta,
                        const int8_t* s_old, int64_t ns, const int8_t* s_new,
                        const int32_t* segment, int64_t nconn, int nflips,
                        double* logabs_out, qtx_stream_t stream);

/* Per-sample log-derivatives O[s, k] = d log psi(s) / d theta_k (Variational.jacobian,
 * variational.py:424-511): O[s, i*N+j] = tanh(theta_i) s_j, O[s, M*N+i] = tanh(theta_i),
 * computed in model dtype and cast to out_dtype (variational.py:491).
 * If col_mean (float64 [M*N+M]) is given the row is centred, and if row_scale (float64 [ns])
 * is given it is scaled: out = (O - mean) * scale, i.e. Obar of quantax/optimizer/sr.py:74-88
 * written in a single pass.  out [ns, ld] with ld >= M*N+M elements per row.
 * tanh_table (nullable): tanh(theta) [ns, M] in model dtype as written by
 * qtx_rbm_jacobian_colmean(tanh_out) for the same spins; skips the theta recomputation. */
int qtx_rbm_jacobian(int model_dtype, const void* W, const void* b, int N, int M,
                     const int8_t* spins, int64_t ns, int out_dtype, void* out, int64_t ld,
                     const double* col_mean, const double* row_scale, const void* tanh_table,
                     qtx_stream_t stream);

/* Column mean of the RBM Jacobian without materialising it (mean over samples of
 * tanh(theta_i) s_j, optionally weighted by w[s]): the `jnp.mean(Omat, axis=0)` of
 * sr.py:77,86.  mean_out float64 [M*N+M]; scratch via qtx_rbm_colmean_workspace_size. */
size_t qtx_rbm_colmean_workspace_size(int model_dtype, int N, int M, int64_t ns);
int qtx_rbm_jacobian_colmean(int model_dtype, const void* W, const void* b, int N, int M,
                             const int8_t* spins, int64_t ns, const double* weight,
                             double* mean_out, void* tanh_out /* nullable, [ns, M] model dtype */,
                             void* workspace, size_t workspace_bytes, qtx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ResConv (quantax/model/conv_nets.py:26-183): pre-activation residual CNN with circular padding,
 * tanh-GELU, final exp / sinh+1 "by scale" activation, channel mean and translation sum.
 *   params: flat vector in the reference's ravel_pytree order (per block conv1.weight
 *   [C,Cin,kh,kw], conv1.bias [C], conv2.weight [C,C,kh,kw], conv2.bias [C]; the last conv has no
 *   bias).  Chains (1-D lattices) use lx = 1, kh = 1.  final_act: 0 = exp_by_scale,
 *   1 = sinhp1_by_scale (quantax/nn/activation.py:7-32).
 *   psi = significand * exp(exponent) (ScaleArray), float64 [ns] each
 *   (quantax/state/variational.py:262-266 with the Identity symmetry).
 * ------------------------------------------------------------------------------------------ */
int64_t qtx_resconv_nparams(int nblocks, int channels, int lx, int ly, int kh, int kw);
/* 1 if the float32 tensor-core tower (csrc/resconv_tc.cu) serves this shape, else 0 (CUDA-core path). */
int qtx_resconv_tc_available(int model_dtype, int channels, int lx, int ly, int kh, int kw);
/* 1 if qtx_resconv_jacobian / _cplx also run their backward pass (variational.py:429-491) on the tensor cores for
 * this shape (wherever the CTA-pair tower serves the forward), else 0 (CUDA-core backward). */
int qtx_resconv_tc_backward_available(int model_dtype, int channels, int lx, int ly, int kh, int kw);
size_t qtx_resconv_workspace_size(int model_dtype, int64_t ns, int nblocks, int channels, int lx,
                                  int ly, int kh, int kw, int need_grad);
/* Batched direct forward: Variational.__call__ / _fulljit_forward (variational.py:268-274,325-347). */
int qtx_resconv_forward(int model_dtype, const void* params, int nblocks, int channels, int lx,
                        int ly, int kh, int kw, int final_act, const int8_t* spins, int64_t ns,
                        double* significand_out, double* exponent_out, void* workspace,
                        size_t workspace_bytes, qtx_stream_t stream);
/* Per-sample log-derivatives (Variational.jacobian, variational.py:424-511): softmax-weighted
 * backprop through the net, exponent is stop-gradient; out [ns, ld] of out_dtype, columns in
 * parameter order.  significand_out / exponent_out are nullable. */
int qtx_resconv_jacobian(int model_dtype, const void* params, int nblocks, int channels, int lx,
                         int ly, int kh, int kw, int final_act, const int8_t* spins, int64_t ns,
                         int out_dtype, void* out, int64_t ld, double* significand_out,
                         double* exponent_out, void* workspace, size_t workspace_bytes,
                         qtx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Model-agnostic Metropolis step for states without local updates (full forward per proposal,
 * variational.py:383-384): propose -> (caller evaluates psi(new_spins)) -> accept.
 * The Philox stream is the one of qtx_rbm_sweep: counter (chain0+chain, step), key seed.
 *   propose: new_spins [ns, N], moved uint8 [ns] (any(s' != s), metropolis.py:314); inj_pos /
 *            inj_slot int32 [ns] for THIS step (nullable).
 *   accept : metropolis.py:299-322 on (mult, expo) pairs; spins / mult / expo updated in place;
 *            inj_u float64 [ns] for this step (nullable); naccept int32 [ns] incremented
 *            (nullable); accept_log uint8 [ns] (nullable).
 * ------------------------------------------------------------------------------------------ */
int qtx_metropolis_propose(int kind, const int8_t* spins, int64_t ns, int N,
                           const int32_t* nbr_table, int max_nb, int hop, const int32_t* inj_pos,
                           const int32_t* inj_slot, uint64_t seed, uint64_t step, uint64_t chain0,
                           int8_t* new_spins, uint8_t* moved, qtx_stream_t stream);
int qtx_metropolis_accept(int8_t* spins, const int8_t* new_spins, const uint8_t* moved, int64_t ns,
                          int N, double* mult, double* expo, const double* mult_new,
                          const double* expo_new, double reweight, const double* inj_u,
                          uint64_t seed, uint64_t step, uint64_t chain0, int32_t* naccept,
                          uint8_t* accept_log, qtx_stream_t stream);

/* Moved proposals only.  The reference evaluates psi(s') of every proposal, including the no-op
 * exchanges of equal spins that `updated = any(s' != s)` can never accept (metropolis.py:262-275,
 * 314-316); here they are skipped:
 *   compact_moved : rank_out int32 [ns] (index among the moved chains, -1 otherwise), the moved rows
 *                   of new_spins gathered into compact_spins_out, the count in count_out int64 [1]
 *   forward_n     : qtx_resconv_forward[_cplx] of the first *ns_dev <= ns_max samples, the count
 *                   read on the device (float32 tensor-core towers; QTX_ERR_UNSUPPORTED otherwise)
 *   accept_compact: qtx_metropolis_accept[_cplx] with psi_new of chain c at index rank[c] */
int qtx_compact_moved(const uint8_t* moved, const int8_t* new_spins, int64_t ns, int N,
                      int32_t* rank_out, int8_t* compact_spins_out, int64_t* count_out,
                      qtx_stream_t stream);
int qtx_resconv_forward_n(int model_dtype, const void* params, int nblocks, int channels, int lx,
                          int ly, int kh, int kw, int final_act, int out_complex,
                          const int8_t* spins, int64_t ns_max, const int64_t* ns_dev,
                          double* significand_out, double* exponent_out, void* workspace,
                          size_t workspace_bytes, qtx_stream_t stream);
int qtx_metropolis_accept_compact(int8_t* spins, const int8_t* new_spins, const uint8_t* moved,
                                  const int32_t* rank, int64_t ns, int N, double* mult, double* expo,
                                  const double* mult_new, const double* expo_new, int mult_complex,
                                  double reweight, const double* inj_u, uint64_t seed, uint64_t step,
                                  uint64_t chain0, int32_t* naccept, uint8_t* accept_log,
                                  qtx_stream_t stream);

/* Whole sweep of a bare ResConv state in ONE call (Metropolis._partial_sweep, metropolis.py:246-275): `nsweeps` times
 * propose -> forward of the (moved) proposals -> accept, enqueued back to back with no host synchronisation; spins
 * [ns, N] in/out, significand / exponent [ns] = psi of the final chains, naccept int32 [ns] nullable.  Same Philox
 * stream, same chains as the step-by-step entry points above.  Workspace from qtx_resconv_sweep_workspace_size. */
size_t qtx_resconv_sweep_workspace_size(int model_dtype, int64_t ns, int nblocks, int channels, int lx, int ly,
                                        int kh, int kw);
int qtx_resconv_sweep(int model_dtype, const void* params, int nblocks, int channels, int lx, int ly, int kh,
                      int kw, int final_act, int8_t* spins, int64_t ns, int nsweeps, int kind,
                      const int32_t* nbr_table, int max_nb, int hop, double reweight, uint64_t seed,
                      uint64_t step0, uint64_t chain0, double* significand_out, double* exponent_out,
                      int32_t* naccept_out, void* workspace, size_t workspace_bytes, qtx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * State-level symmetry projection psi(s) = sum_g w_g psi(T_g s), w_g = chi_g chi_0 / |G|
 * (quantax/state/variational.py:262-266, quantax/symmetry/symmetry.py:325-392).
 *   images : out int8 [ns, nsymm, N], s_g = s[perm_g]; with z2 != 0 the second half is -s[perm_g]
 *            (nsymm = 2 nperm).  perm int32 [nperm, N].
 *   combine: signed log-sum-exp over the images in container arithmetic; (mult, expo) [ns, nsymm]
 *            -> [ns].  kind 0 = LogArray result (sign, logabs), kind 1 = ScaleArray result
 *            (significand, exponent).  coef_out (nullable) float64 [ns, nsymm] = w_g psi_g / psi,
 *            the image weights of the projected log-derivative (variational.py:438-491).
 *   weighted_rowsum: out[s, :] = sum_g coef[s, g] J[s nsymm + g, :]  (projected Jacobian rows).
 * ------------------------------------------------------------------------------------------ */
int qtx_symm_images(const int8_t* spins, int64_t ns, int N, const int32_t* perm, int nperm, int z2,
                    int8_t* out, qtx_stream_t stream);
int qtx_symm_combine(const double* mult, const double* expo, int64_t ns, int nsymm,
                     const double* weights, int kind, double* mult_out, double* expo_out,
                     double* coef_out, qtx_stream_t stream);
int qtx_weighted_rowsum(int dtype, const void* J, int64_t ldj, const double* coef, int64_t ns,
                        int nsymm, int64_t np, void* out, int64_t ldo, qtx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Connected-configuration enumeration for generic models (bit-exact mirror of
 * _apply_off_diag + _get_conn, operator.py:96-165).  Terms with `nflips_sel` flips only.
 *   count:  nonnan_out / valid_out int32 [ns]  (valid = not NaN and |H| > 1e-8)
 *   fill:   offsets int64 [ns] = exclusive scan of valid counts (per device range);
 *           writes segment int32, conn_idx int32, H float64 [conn_size] and (nullable)
 *           s_conn int8 [conn_size, N]; entries past the total are padding
 *           (segment -1, conn_idx -1, H 0, s_conn = last raw candidate of the last sample).
 * ------------------------------------------------------------------------------------------ */
int qtx_conn_count(const int8_t* spins, int64_t ns, int N, const double* term_coef,
                   const uint16_t* term_sites, const uint8_t* term_ops, int nterms,
                   int nflips_sel, int32_t* nonnan_out, int32_t* valid_out, qtx_stream_t stream);
int qtx_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, int64_t* total_out,
                           qtx_stream_t stream);
int qtx_conn_fill(const int8_t* spins, int64_t ns, int N, const double* term_coef,
                  const uint16_t* term_sites, const uint8_t* term_ops, int nterms,
                  int nflips_sel, const int64_t* offsets, const int64_t* total, int64_t conn_size,
                  int32_t* segment_out,
                  int32_t* conn_idx_out, double* H_out, int8_t* s_conn_out, qtx_stream_t stream);
/* diagonal part sum_t J_t prod_k (s_k / 2) over all-diagonal terms (operator.py:81-93) */
int qtx_apply_diag(const int8_t* spins, int64_t ns, int N, const double* term_coef,
                   const uint16_t* term_sites, const uint8_t* term_ops, int nterms,
                   double* diag_out, qtx_stream_t stream);
/* Eloc[seg] += H * (mult'/mult[seg]) * exp(expo' - expo[seg])  (_get_Olocx, operator.py:168-184) */
int qtx_oloc_reduce(const int32_t* segment, const double* H, const double* mult_conn,
                    const double* expo_conn, int64_t nconn, const double* mult, const double* expo,
                    int64_t ns, double* eloc_inout, qtx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SR / MinSR dense algebra (quantax/optimizer/sr.py:74-123, solver.py:94-201)
 * ------------------------------------------------------------------------------------------ */
/* mean_out[k] = (1/ns) sum_s w[s] A[s,k]  (w nullable = 1);  A [ns, ld] of dtype */
int qtx_colmean(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, const double* weight,
                double* mean_out, qtx_stream_t stream);
/* A[s,k] = (A[s,k] - mean[k]) * scale[s] in place (_Omat_to_Obar, sr.py:74-77) */
int qtx_center_scale(int dtype, void* A, int64_t ns, int64_t np, int64_t ld, const double* mean,
                     const double* scale, qtx_stream_t stream);
/* Ebar, energy, VarE from local energies (SR.get_Ebar, sr.py:180-195); stats_out float64 [2]
 * = {energy, VarE} on device. */
int qtx_ebar(const double* eloc, const double* rw, int64_t ns, double* ebar_out, double* stats_out,
             qtx_stream_t stream);

/* T = A A^T  (solver.py:139), A [ns, ld] (np used columns), T_out float64/float32 [ns, ns]
 * (both triangles written).  The contraction runs on tcgen05 tensor cores fed by TMA:
 *   QTX_F64 input: error-free int8 slicing (Ozaki scheme) with `nslices` 7-bit slices
 *   (0 = default for the dtype), exact int32 accumulation, float64 recombination;
 *   QTX_F32 input: same with fewer slices.
 * `T_accum != 0` adds into T_out (used to sum column shards). */
size_t qtx_gram_workspace_size(int dtype, int64_t ns, int64_t np, int nslices);
int qtx_gram(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices,
             double* T_out, int T_accum, void* workspace, size_t workspace_bytes,
             qtx_stream_t stream);

/* ---- fused Gram + exchange over peer memory (distributed MinSR, solver.py:134-139) --------------
 * The reference sums the partial Gram matrices of the column shards through XLA's GSPMD
 * all-reduce.  Here every rank owns a staging area stage[P][ns][ns] (float64) and a flag array
 * flags[P] (uint64), allocated with qtx_peer_alloc, exported with qtx_peer_export (64-byte CUDA
 * IPC handle) and mapped by the other ranks of the node with qtx_peer_open.
 *   qtx_gram_push   : qtx_gram whose epilogue also stores every finished tile (j <= i) into
 *                     peer_slots[q] = stage_q + rank * ns * ns for all q != rank (NVLink stores
 *                     overlapped with the MMAs of the next tile);
 *   qtx_peer_signal : flags_q[rank] = epoch on every rank q (release at system scope), enqueued
 *                     after the push;
 *   qtx_gram_reduce : waits (bounded by timeout_s, default 60 s; traps on expiry) until
 *                     my_flags[q] >= epoch for all q, then T_out[i,j] = T_out[j,i] =
 *                     sum_q partials[q][i,j] in rank order (bit-identical on all ranks);
 *                     partials[rank] may alias T_out.  my_flags = NULL skips the wait.
 * Host pointer arrays (peer_slots, peer_flags, partials) hold nranks device pointers. */
int qtx_peer_alloc(size_t bytes, void** ptr_out);
int qtx_peer_free(void* ptr);
int qtx_peer_export(void* ptr, void* handle64_out);
int qtx_peer_open(const void* handle64, void** ptr_out);
int qtx_peer_close(void* ptr);
int qtx_gram_push(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, int nslices,
                  double* T_out, int nranks, int rank, void* const* peer_slots, void* workspace,
                  size_t workspace_bytes, qtx_stream_t stream);
int qtx_peer_signal(void* const* peer_flags, int nranks, int rank, uint64_t epoch,
                    qtx_stream_t stream);
int qtx_gram_reduce(const void* const* partials, int nranks, int64_t ns, double* T_out,
                    const void* my_flags, uint64_t epoch, double timeout_s, qtx_stream_t stream);

/* (lambda, U) = eigh(T); y = U (lambda^+ o (U^T b)) with the soft pseudo-inverse
 * lambda^+ = 1 / (lambda (1 + ((rtol max|lambda| + atol)/|lambda|)^6)), 0 where lambda == 0
 * (solver.py:94-101,142-146; minsr_pinv_eig solver.py:262-294).  rtol < 0 selects the dtype
 * default (1e-12).  T [n, n] float64 is overwritten by U (column-major eigenvectors =
 * row-major U^T); evals_out float64 [n] (nullable); y_out float64 [n].
 * info_out int32 [1] device (0 = converged).  eigh is cuSOLVER syevd (library call). */
size_t qtx_pinv_eig_workspace_size(int64_t n);
int qtx_pinv_eig_solve(double* T, int64_t n, const double* b, double rtol, double atol,
                       double* evals_out, double* y_out, int32_t* info_out, void* workspace,
                       size_t workspace_bytes, qtx_stream_t stream);

/* Same with the signal-to-noise damping of `_sum_without_noise` (solver.py:114-125): rho_k = U_k^T b is
 * divided by 1 + (tol_snr / snr_k)^6, snr_k = |mean_t r_tk| / sqrt(mean_t |r_tk - mean|^2 / n),
 * r_tk = U[t,k] b[t].  tol_snr <= 1e-6 is the plain sum (the reference's `cond`). */
int qtx_pinv_eig_solve_snr(double* T, int64_t n, const double* b, double rtol, double atol,
                           double tol_snr, double* evals_out, double* y_out, int32_t* info_out,
                           void* workspace, size_t workspace_bytes, qtx_stream_t stream);
/* The three stages separately, for solvers whose rho is not U^T b (lstsq_pinv_eig with tol_snr,
 * solver.py:156-162: rho_sk = (A V)[s,k] b[s]).  qtx_eigh: T -> row-major U^T, evals_out [n]
 * (workspace of qtx_pinv_eig_workspace_size).  qtx_rows_dot_snr: rho[k] = sum_without_noise_i
 * (M[k,i] b[i]) for M [nrows, ld].  qtx_pinv_apply: y = U (lambda^+ o rho); rho is overwritten
 * by lambda^+ o rho. */
int qtx_eigh(double* T, int64_t n, double* evals_out, int32_t* info_out, void* workspace,
             size_t workspace_bytes, qtx_stream_t stream);
int qtx_rows_dot_snr(const double* M, int64_t nrows, int64_t n, int64_t ld, const double* b,
                     double tol_snr, double* rho_out, qtx_stream_t stream);
int qtx_pinv_apply(const double* Ut, int64_t n, const double* evals, double* rho_inout, double rtol,
                   double atol, double* y_out, qtx_stream_t stream);

/* The same soft pseudo-inverse y = f(T) b, f(lambda) = lambda^5 / (lambda^6 + c^6), c = rtol max|lambda| + atol
 * (solver.py:94-111,142-146 after `eigh`), WITHOUT an eigendecomposition: by partial fractions over the roots
 * z_k = c exp(i pi (2k+1)/6) of lambda^6 + c^6,  f(T) b = (1/3) Re sum_{k=0,1,2} (T - z_k I)^-1 b  exactly.
 *   qtx_sym_absmax_eig        : lam_out [1] device = max|lambda| of the symmetric T [n, n] from `steps` Lanczos
 *                               steps (three-term recurrence + bisection; steps <= 1024).  first_step = 0 starts
 *                               the recurrence; first_step = the `steps` of the previous call continues it from
 *                               the state kept in the (untouched) workspace, so a caller can double the number
 *                               of steps until the value settles.
 *   qtx_pinv_rational_partial : for every k in shift_mask (bit k), complex LU of T - z_k I (cuSOLVER Zgetrf /
 *                               Zgetrs, library calls), `refine_steps` refinement steps with the residual in
 *                               double-double arithmetic, and ydd (+)= Re x_k as a double-double vector
 *                               ydd_inout float64 [2][n] = (hi, lo).  T is NOT overwritten.  The ranks of a
 *                               distributed solve take different shifts.  rtol < 0 selects 1e-12; rtol = atol = 0
 *                               is QTX_ERR_UNSUPPORTED (plain inverse: use qtx_pinv_eig_solve).
 *                               info_out int32 [1] device (0 = all factorizations succeeded).
 *   qtx_dd_sum_scale          : y_out [n] = scale * sum_q ydd[q] for `count` double-double vectors
 *                               ydd float64 [count][2][n], summed in order and rounded once (scale = 1/3).
 * All three use the workspace of qtx_pinv_rational_workspace_size (0 on failure). */
size_t qtx_pinv_rational_workspace_size(int64_t n);
int qtx_sym_absmax_eig(const double* T, int64_t n, int first_step, int steps, double* lam_out,
                       void* workspace, size_t workspace_bytes, qtx_stream_t stream);
int qtx_pinv_rational_partial(const double* T, int64_t n, const double* b, double rtol, double atol,
                              const double* lam, int shift_mask, int refine_steps, double* ydd_inout,
                              int accumulate, int32_t* info_out, void* workspace, size_t workspace_bytes,
                              qtx_stream_t stream);
int qtx_dd_sum_scale(const double* ydd, int count, int64_t n, double scale, double* y_out,
                     qtx_stream_t stream);

/* The same partial sums with the library's OWN kernels and no cuSOLVER call -- the default soft pseudo-inverse of
 * the SR / MinSR solve (replaces `eigh` + `_get_eigs_inv`, solver.py:94-111,142-146): T - z_k I is complex
 * symmetric with its field of values off the origin, so it is factorised as L D L^T WITHOUT pivoting (csrc/zldlt.cu:
 * blocked right-looking, FP64 FMA trailing updates, wavefront triangular solves), refined with double-double
 * residuals like qtx_pinv_rational_partial.  Several shifts of one call run concurrently on side streams that are
 * forked from and joined back into `stream`.  info_out: 0 = ok, > 0 = 1-based index of a zero pivot, < 0 = -(k+1):
 * the refinement of shift k did not contract.
 *   qtx_pinv_ldlt_workspace_size(n, nshifts) : bytes for a call that takes `nshifts` (1..3) shifts (0 on failure)
 *   qtx_sym_absmax_eig_ws                    : qtx_sym_absmax_eig inside that workspace (same continuation protocol) */
size_t qtx_pinv_ldlt_workspace_size(int64_t n, int nshifts);
int qtx_sym_absmax_eig_ws(const double* T, int64_t n, int first_step, int steps, double* lam_out,
                          void* workspace, size_t workspace_bytes, int nshifts, qtx_stream_t stream);
int qtx_pinv_ldlt_partial(const double* T, int64_t n, const double* b, double rtol, double atol,
                          const double* lam, int shift_mask, int refine_steps, double* ydd_inout,
                          int accumulate, int32_t* info_out, void* workspace, size_t workspace_bytes,
                          qtx_stream_t stream);

/* y = (T + shift I)^-1 b, shift = rshift * trace(T) + ashift, by Cholesky (minnorm_shift_eig /
 * lstsq_shift_eig, solver.py:50-77; `solve(assume_a="pos")`).  rshift < 0 selects the dtype
 * default (1e-12).  T [n, n] float64 symmetric is overwritten by its factor; info_out int32 [1]
 * device (0 = positive definite).  potrf / potrs are cuSOLVER library calls. */
size_t qtx_shift_chol_workspace_size(int64_t n);
int qtx_shift_chol_solve(double* T, int64_t n, const double* b, double rshift, double ashift,
                         double* y_out, int32_t* info_out, void* workspace, size_t workspace_bytes,
                         qtx_stream_t stream);

/* out[k] = sum_s A[s,k]^2  (diagonal of S = A^T A in lstsq_shift_cg, solver.py:35) */
int qtx_col_sumsq(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, double* out,
                  qtx_stream_t stream);

/* x[k] = sum_s A[s,k] y[s]  (the final A^dagger y of solver.py:146); x_out float64 [np] */
int qtx_matvec_t(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, const double* y,
                 double* x_out, int accumulate, qtx_stream_t stream);
/* v[s] = sum_k A[s,k] x[k]; v_out float64 [ns] */
int qtx_matvec(int dtype, const void* A, int64_t ns, int64_t np, int64_t ld, const double* x,
               double* v_out, qtx_stream_t stream);
/* params <- params - step (Variational.update, variational.py:558-579); the update is skipped
 * when any step entry is non-finite; flag_out int32 [1] = 1 if applied. */
int qtx_apply_update(int model_dtype, void* params, const double* step, double lr, int64_t np,
                     int32_t* flag_out, qtx_stream_t stream);

/* Vector helpers of the momentum optimizers SPRING / MARCH / AdamSR
 * (quantax/optimizer/sr.py:198-429), all float64 device vectors:
 *   axpby          y <- a x + b y
 *   div_add        out <- x / d + c z            (z nullable)
 *   second_moment  V <- beta V + (1 - beta) |x - y|^2   (y nullable)
 *   fourth_root    out <- (v / corr)^(1/4) + eps
 *   scale_columns  A[s, k] <- A[s, k] / d[k]     (Obar /= V[None, :], sr.py:304,409) */
int qtx_axpby(int64_t n, double a, const double* x, double b, double* y, qtx_stream_t stream);
/* S[i, j] += alpha x[i] x[j], S float64 [n, n]  (TimeEvol, quantax/optimizer/time_evol.py:113-114) */
int qtx_rank1_update(int64_t n, double alpha, const double* x, double* S, qtx_stream_t stream);
int qtx_div_add(int64_t n, const double* x, const double* d, double c, const double* z, double* out,
                qtx_stream_t stream);
int qtx_second_moment(int64_t n, double beta, const double* x, const double* y, double* V,
                      qtx_stream_t stream);
int qtx_fourth_root(int64_t n, const double* v, double corr, double eps, double* out,
                    qtx_stream_t stream);
int qtx_scale_columns(int dtype, void* A, int64_t ns, int64_t np, int64_t ld, const double* d,
                      qtx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Complex-output states with real parameters (VS_TYPE.real_to_complex,
 * quantax/state/variational.py:244-257): ResConv with out_dtype complex (pair_cpl,
 * quantax/model/conv_nets.py:165-170, quantax/nn/activation.py:75-81), sign / phase layers
 * (quantax/nn/sign.py:8-75), complex Oloc, symmetry projection and the stacked [Re; Im] rows that
 * QNGD.solve builds for the real solver (quantax/optimizer/sr.py:99-104).
 * Complex vectors are complex128, interleaved (re, im): "c128" in the parameter name.
 * ------------------------------------------------------------------------------------------ */
/* psi = ScaleArray(significand complex128 [ns], exponent float64 [ns]); channels must be even. */
int qtx_resconv_forward_cplx(int model_dtype, const void* params, int nblocks, int channels, int lx,
                             int ly, int kh, int kw, int final_act, const int8_t* spins, int64_t ns,
                             double* significand_c128_out, double* exponent_out, void* workspace,
                             size_t workspace_bytes, qtx_stream_t stream);
/* out rows [0, ns) = Re O, rows [im_row_offset, im_row_offset + ns) = Im O  (variational.py:461-487:
 * two backward passes seeded with d Re(log psi) and d Im(log psi)). */
int qtx_resconv_jacobian_cplx(int model_dtype, const void* params, int nblocks, int channels, int lx,
                              int ly, int kh, int kw, int final_act, const int8_t* spins, int64_t ns,
                              int out_dtype, void* out, int64_t ld, int64_t im_row_offset,
                              double* significand_c128_out, double* exponent_out, void* workspace,
                              size_t workspace_bytes, qtx_stream_t stream);
/* mult[s] *= exp(i * dot(kernel, s)), float32 dot product (sign.py:31,36: compute_sign "phase"). */
int qtx_apply_sign_phase(const float* kernel, const int8_t* spins, int64_t ns, int N,
                         double* mult_c128, qtx_stream_t stream);
/* qtx_metropolis_accept with complex128 mult / mult_new. */
int qtx_metropolis_accept_cplx(int8_t* spins, const int8_t* new_spins, const uint8_t* moved,
                               int64_t ns, int N, double* mult_c128, double* expo,
                               const double* mult_new_c128, const double* expo_new, double reweight,
                               const double* inj_u, uint64_t seed, uint64_t step, uint64_t chain0,
                               int32_t* naccept, uint8_t* accept_log, qtx_stream_t stream);
/* qtx_oloc_reduce with complex128 amplitudes; eloc complex128 [ns], accumulated into. */
int qtx_oloc_reduce_cplx(const int32_t* segment, const double* H, const double* mult_conn_c128,
                         const double* expo_conn, int64_t nconn, const double* mult_c128,
                         const double* expo, int64_t ns, double* eloc_c128_inout,
                         qtx_stream_t stream);
/* qtx_symm_combine for ScaleArray images with complex significands (real weights);
 * coef_c128_out [ns, nsymm] nullable. */
int qtx_symm_combine_cplx(const double* mult_c128, const double* expo, int64_t ns, int nsymm,
                          const double* weights, double* mult_c128_out, double* expo_out,
                          double* coef_c128_out, qtx_stream_t stream);
/* Projected Jacobian of a complex-output state on stacked real matrices: J rows [0, ns*nsymm) = Re,
 * [j_im_row_offset, ...) = Im of O(T_g s); out rows [0, ns) = Re, [out_im_row_offset, ...) = Im. */
int qtx_weighted_rowsum_cplx(int dtype, const void* J, int64_t ldj, int64_t j_im_row_offset,
                             const double* coef_c128, int64_t ns, int nsymm, int64_t np, void* out,
                             int64_t ldo, int64_t out_im_row_offset, qtx_stream_t stream);
/* SR.get_Ebar for complex local energies: stats = (Re <E rw>, <|E - <E rw>|^2 rw>); ebar stacked
 * float64: [s] = Re, [im_offset + s] = Im of (E - <E>) sqrt(rw / ns)  (sr.py:102,180-195). */
int qtx_ebar_cplx(const double* eloc_c128, const double* rw, int64_t ns, double* ebar_stacked_out,
                  int64_t im_offset, double* stats_out, qtx_stream_t stream);
/* ------------------------------------------------------------------------------------------
 * RBM_Conv / SingleConv (quantax/model/shallow_nets.py:129-190): one full-lattice circular
 * convolution followed by prod cosh = a dense RBM with M = channels * N tied hidden units.
 *   expand   : kernel [channels, lx*ly] (Conv.weight [C,1,Lx,Ly]), bias [channels] (nullable) ->
 *              W [M, N] row-major and b [M] for the qtx_rbm_* entry points
 *   jacobian : theta [ns, M] (qtx_rbm_forward) -> out[s, c*N + d] = sum_r tanh(theta_{c,r}) s[(r+d-lo) mod L],
 *              out[s, channels*N + c] = sum_r tanh(theta_{c,r})  (parameter order: weight, bias)
 * ------------------------------------------------------------------------------------------ */
int qtx_rbm_conv_expand(int model_dtype, const void* kernel, const void* bias, int channels, int lx,
                        int ly, void* W_out, void* b_out, qtx_stream_t stream);
int qtx_rbm_conv_jacobian(int model_dtype, const void* theta, const int8_t* spins, int64_t ns,
                          int channels, int lx, int ly, int out_dtype, void* out, int64_t ld,
                          qtx_stream_t stream);
/* out[i] = x[i] + 0i */
int qtx_real_to_cplx(const double* x, int64_t n, double* out_c128, qtx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Collectives of the data-parallel step: what the reference gets implicitly from GSPMD on its
 * sharded arrays (quantax/utils/function.py:163-214, quantax/optimizer/solver.py:134-139,146) as explicit calls
 * on an NCCL communicator, so that any host -- the torch binding of this repository, a jax.ffi binder -- can run
 * the multi-GPU step through this ABI alone.  NCCL is bound at run time (dlopen libnccl.so.2); without it these
 * entry points return QTX_ERR_UNSUPPORTED and everything else works.
 *   qtx_comm_unique_id : 128-byte ncclUniqueId (rank 0 creates it, the host broadcasts it by its own means)
 *   qtx_comm_init      : ncclCommInitRank; qtx_comm_adopt wraps an ncclComm_t the host already owns
 *   qtx_comm_all_reduce: in-place allowed; dtype QTX_F32 / QTX_F64 / QTX_I32, op QTX_SUM / QTX_MAX
 *   qtx_comm_all_gather: recv [nranks][bytes_per_rank]
 *   qtx_comm_all_to_all: recv block p <- send block `rank` of rank p (grouped ncclSend / ncclRecv)
 * All enqueue on `stream`.
 * ------------------------------------------------------------------------------------------ */
int qtx_comm_unique_id(void* id_out_128_bytes);
int qtx_comm_init(qtx_comm_t* comm_out, int nranks, int rank, const void* id_128_bytes);
int qtx_comm_adopt(qtx_comm_t* comm_out, void* nccl_comm);
int qtx_comm_destroy(qtx_comm_t comm);
int qtx_comm_size(qtx_comm_t comm);
int qtx_comm_rank(qtx_comm_t comm);
int qtx_comm_all_reduce(qtx_comm_t comm, const void* send, void* recv, int64_t count, int dtype, int op,
                        qtx_stream_t stream);
int qtx_comm_all_gather(qtx_comm_t comm, const void* send, void* recv, int64_t bytes_per_rank,
                        qtx_stream_t stream);
int qtx_comm_broadcast(qtx_comm_t comm, void* buf, int64_t bytes, int root, qtx_stream_t stream);
int qtx_comm_all_to_all(qtx_comm_t comm, const void* send, void* recv, int64_t bytes_per_peer,
                        qtx_stream_t stream);

/* Distributed MinSR solve x = A^+ b with the ROWS of A = Obar sharded over the ranks of `comm`
 * (minnorm_pinv_eig under GSPMD, solver.py:128-149): column shards by all-to-all (parameter axis zero-padded to a
 * multiple of nranks, solver.py:136), tensor-core Gram of the shard, all-reduce of T (solver.py:139), the soft
 * pseudo-inverse y = f(T) b with the three shifted LDL^T solves split over the ranks (qtx_pinv_ldlt_partial) and
 * their double-double partial sums all-gathered and added in rank order (bit-identical y everywhere), the column
 * shard of x = A^T y, all-gather (solver.py:146).  A_local [nl, np] (ld), b_local [nl], x_out [np] float64 on every
 * rank, info_out int32 [1] device (0 = ok).  max|lambda|: `lanczos_steps` > 0 = adaptive Lanczos run of at most that
 * many steps (stages 32, 64, 128, ... until two stages agree to 1e-7, one scalar read-back per stage, the rule of the
 * single-GPU solve; all ranks stop at the same stage because T is bit-identical); < 0 = exactly |lanczos_steps| steps
 * without any host synchronisation (128 suffices unless the top of the spectrum is dense). */
size_t qtx_minsr_solve_dist_workspace_size(qtx_comm_t comm, int dtype, int64_t nl, int64_t np, int nslices);
int qtx_minsr_solve_dist(qtx_comm_t comm, int dtype, const void* A_local, int64_t nl, int64_t np, int64_t ld,
                         const double* b_local, double rtol, double atol, int nslices, int lanczos_steps,
                         int refine_steps, double* x_out, int32_t* info_out, void* workspace,
                         size_t workspace_bytes, qtx_stream_t stream);
/* Phase timing of qtx_minsr_solve_dist (measurement aid, off by default): after qtx_minsr_solve_dist_timing(1) every
 * solve records CUDA events on its stream; qtx_minsr_solve_dist_phases waits for the last solve and returns the
 * milliseconds of  [0] pack + all-to-all of Obar  [1] Gram of the column shard  [2] all-reduce of T + all-gather of b
 * [3] Lanczos max|lambda|  [4] shifted LDL^T solves of this rank  [5] all-gather + rank-ordered sum of y
 * [6] column shard of A^T y  [7] all-gather of x. */
int qtx_minsr_solve_dist_timing(int enable);
int qtx_minsr_solve_dist_phases(double* ms_out_8);

#ifdef __cplusplus
}
#endif
#endif /* QTX_B200_H */
