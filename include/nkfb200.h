/*
 * nkfb200 -- C ABI of the B200-native (sm_100a) hot path of netket_fidelity:
 * one infidelity-gradient step of p-tVMC.
 *
 * Every entry point is `extern "C"`, takes plain pointers/sizes and a
 * `cudaStream_t` (passed as void*), returns 0 on success or a negative
 * NKFB_ERR_* code, never throws and never calls exit().  Device memory is owned
 * by the caller (the Python host passes `tensor.data_ptr()`); the library only
 * reads/writes through those pointers on `stream`.  The only library-owned
 * memory is the per-handle workspace used by the *_host entry points.
 *
 * Each declaration cites the reference interface it replaces
 * (paths relative to the netket_fidelity repository).
 *
 * Data layouts
 *  - configurations ("samples") are bit-packed: uint32 words, W32 = ceil(N/32)
 *    words per configuration, site i -> bit (i & 31) of word (i >> 5);
 *    bit 0 <-> local_states[0], bit 1 <-> local_states[1]
 *    (Spin(1/2): -1 / +1, Qubit: 0 / 1).
 *  - RBM parameters are complex, interleaved (re, im), in the real type named
 *    by `dtype` (NKFB_F32 -> complex64, NKFB_F64 -> complex128):
 *    kernel (N, M) row-major (the NetKet `Dense/kernel` leaf), bias (M),
 *    visible_bias (N).  Real-parameter RBMs are passed with zero imaginary
 *    parts; their gradient is the real part of the complex-convention one.
 *  - flat gradient / accumulator layout: [bias (M) | kernel (N*M) | visible_bias (N)],
 *    complex, i.e. the leaf order of jax.tree_util on the NetKet RBM parameter dict.
 */
#ifndef NKFB200_H
#define NKFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NKFB_VERSION 100

enum {
  NKFB_OK = 0,
  NKFB_ERR_INVALID = -1,     /* bad argument */
  NKFB_ERR_UNSUPPORTED = -2, /* shape / mode outside what the kernels cover */
  NKFB_ERR_CUDA = -3,        /* a CUDA runtime call failed; see nkfb_last_error */
  NKFB_ERR_NOMEM = -4
};

enum { NKFB_F32 = 0, NKFB_F64 = 1 };

enum { NKFB_OP_NONE = 0, NKFB_OP_GATE = 1, NKFB_OP_ISING = 2 };

typedef struct nkfb_handle_s *nkfb_handle_t;

/* RBM parameters (NetKet nk.models.RBM; verbatim form in
 * examples/twospins_GHZ/RBM_Jastrow_ansatz.py:37-62). */
typedef struct {
  int32_t n_visible;        /* N */
  int32_t n_hidden;         /* M = alpha N */
  int32_t dtype;            /* NKFB_F32 | NKFB_F64 */
  int32_t reserved;
  const void *kernel;       /* device, (N, M) complex */
  const void *bias;         /* device, (M) complex, or NULL (use_hidden_bias=False) */
  const void *visible_bias; /* device, (N) complex, or NULL (use_visible_bias=False) */
} nkfb_rbm_t;

/* A discrete operator given by its connected elements.
 *  GATE : any single-site operator (Rx, Ry, Hadamard:
 *         netket_fidelity/operator/singlequbit_gates.py:89-105,184-201,270-288).
 *         conn 0 = x, conn 1 = x with site `idx` flipped;
 *         mel[c] = gate_mel[bit of x at idx][c] (re, im).
 *  ISING: conn 0 = x with mel = c0 + c1 * sum_<ij> (2[x_i==x_j]-1);
 *         conn i+1 = x with site i flipped, mel = c2   (NetKet IsingJax,
 *         netket_fidelity/operator/__init__.py:4; propagator 1 - i dt H likewise). */
typedef struct {
  int32_t kind; /* NKFB_OP_* */
  int32_t idx;
  double gate_mel[2][2][2];
  double c0[2], c1[2], c2[2];
  int32_t n_edges;
  int32_t reserved;
  const int32_t *edges; /* device, (n_edges, 2) */
} nkfb_op_t;

typedef struct {
  int64_t n_chains;      /* chains handled by THIS call (this rank's shard) */
  int32_t chain_length;  /* samples emitted per chain */
  int32_t n_discard;     /* sweeps discarded before the first emitted sample */
  int32_t sweep_size;    /* MH steps per sweep (NetKet default: N) */
  int32_t site_mode;     /* 0: one site sequence shared by all chains; 1: per-chain sites (NetKet) */
  uint64_t seed;
  uint64_t chain_offset; /* global id of this call's first chain (multi-GPU sharding) */
  uint64_t step_offset;  /* MH steps already consumed by earlier calls (Philox counter offset) */
  double local_states[2];
  int32_t init_random;   /* !=0: ignore chain_state on input, start from Philox-random configurations */
  int32_t flags;         /* NKFB_SAMPLE_* bits (0 = none; was `reserved`) */
  void *workspace;         /* device scratch of >= nkfb_sample_workspace_bytes() bytes, or NULL to use the
                              handle's own buffer (then calls on different streams must not overlap) */
  uint64_t workspace_bytes;
  const uint64_t *step_offset_dev; /* optional device counter ADDED to step_offset when the kernels start: lets a
                                      captured CUDA graph of a whole step be replayed (the host-side value is
                                      frozen at capture time); NULL otherwise */
} nkfb_sample_cfg_t;

/* ---- lifecycle ---------------------------------------------------------- */
int nkfb_version(void);
int nkfb_create(int device, nkfb_handle_t *out);
int nkfb_destroy(nkfb_handle_t h);
const char *nkfb_last_error(nkfb_handle_t h);
/* number of kernels this handle has launched so far (bench.py: gpu_launches) */
int64_t nkfb_launch_count(nkfb_handle_t h);

/* Options.  NKFB_OPT_SAMPLER_ALGO selects the plain-target Metropolis kernel of nkfb_sample:
 *   0 (default) automatic: the multiplicative kernel (z = exp(-2 theta) tracked by complex multiplications,
 *               no transcendental per hidden unit) in single precision, the additive one (theta tracked by
 *               additions, EX2 + COS per hidden unit) in double precision;
 *   1           always additive;   2  multiplicative wherever its range windows cover the parameters.
 * Both sample the same chain from the same random numbers (they accept the same proposals up to the rounding
 * of the log-ratio); the additive kernel is always launched behind the multiplicative ones and takes over when
 * max_j (|Re b_j| + max|s| sum_i |Re W_ij|) exceeds their range. */
/* NKFB_OPT_SAMPLER_SHARE (default 1): how many nkfb_sample launches of equal size the caller runs side by side
 * on different streams (2 for the two chain families of an infidelity step).  Only a scheduling hint: it decides
 * whether the chains beyond the last full wave of CTAs are handed to the one-chain-per-warp kernel. */
/* NKFB_OPT_FWD_MAX_CTAS (default 0 = no limit): upper bound on the CTAs of the next nkfb_infidelity_fwd launches.  Lets a
 * caller run pass A on the SMs a concurrent sub-wave sampler launch leaves idle (engine.py, streamed segments). */
/* NKFB_OPT_TC_TABLES_VALID (default 0): 1 = the parameter tables of the tensor-core estimator / gradient kernels (bf16
 * splits of the RBM kernels, written by nkfb_tc_prepare or by the last nkfb_infidelity_fwd / nkfb_infidelity_grad launch
 * into handle-owned scratch) still match the parameter buffers: launches on the SAME buffers and shapes skip their
 * preparation kernels (the pass A launches of the sample segments of one step; pass B after an early nkfb_tc_prepare).
 * The caller clears it (0) before the parameters change; other buffers or shapes are re-prepared whatever the flag says. */
enum { NKFB_OPT_SAMPLER_ALGO = 1, NKFB_OPT_SAMPLER_SHARE = 2, NKFB_OPT_FWD_MAX_CTAS = 3, NKFB_OPT_TC_TABLES_VALID = 4 };
/* nkfb_sample_cfg_t.flags.  NKFB_SAMPLE_TABLES_VALID: `workspace` (caller-provided) still holds the parameter tables an
 * earlier nkfb_sample call wrote for the SAME net, local_states and plain target; their preparation kernels are skipped
 * (a chain sampled in several consecutive calls between two parameter updates).  Ignored for composite targets. */
enum { NKFB_SAMPLE_TABLES_VALID = 1 };
int nkfb_set_option(nkfb_handle_t h, int key, int64_t value);

/* ---- configurations <-> packed bits ------------------------------------- */
/* x: device (B, N) of dtype NKFB_F32/NKFB_F64 holding local-state values. */
int nkfb_pack(nkfb_handle_t h, const void *x, int x_dtype, int64_t B, int N,
              const double local_states[2], uint32_t *packed, void *stream);
int nkfb_unpack(nkfb_handle_t h, const uint32_t *packed, int64_t B, int N,
                const double local_states[2], void *x, int x_dtype, void *stream);

/* ---- connected elements (materialising) --------------------------------- */
/* Replaces Rx/Ry/Hadamard.get_conn_padded (singlequbit_gates.py:65-71,160-166,250-256)
 * and IsingJax.get_conn_padded.  x (B,N) -> xp (B,C,N) same dtype, mels (B,C) complex128.
 * C = 2 (GATE) or N+1 (ISING).  Bit-exact target. */
int nkfb_conn(nkfb_handle_t h, const nkfb_op_t *op, const void *x, int x_dtype,
              int64_t B, int N, const double local_states[2], void *xp, void *mels,
              void *stream);

/* ---- log-amplitudes ------------------------------------------------------ */
/* out[b] = log sum_c mel_c psi(x'_c)  (op == NULL or NONE: log psi(x)).
 * Replaces afun(variables, x) of nk.models.RBM and _logpsi_U_fun
 * (netket_fidelity/utils/sampling_Ustate.py:34-47).  out: (B) complex of net->dtype. */
int nkfb_logpsi(nkfb_handle_t h, const nkfb_rbm_t *net, const nkfb_op_t *op,
                const uint32_t *packed, int64_t B, const double local_states[2],
                void *out, void *stream);

/* ---- Metropolis-local sampler ------------------------------------------- */
/* Replaces nk.sampler.MetropolisLocal + MCState.samples/reset (call sites
 * infidelity/overlap/expect.py:26-27, overlap_U/expect.py:21-22,
 * driver/infidelity_optimizer.py:142-143).  Target density |sum_c mel_c psi(x'_c)|^2.
 * chain_state: device (n_chains, W32) in/out.  samples: device
 * (n_chains, chain_length, W32).  n_accepted: device uint64[1] (accumulated) or NULL. */
size_t nkfb_sample_workspace_bytes(int n_visible, int n_hidden, int dtype);
int nkfb_sample(nkfb_handle_t h, const nkfb_rbm_t *net, const nkfb_op_t *op,
                const nkfb_sample_cfg_t *cfg, uint32_t *chain_state, uint32_t *samples,
                unsigned long long *n_accepted, void *stream);

/* ---- infidelity estimator (pass A) -------------------------------------- */
/* Replaces kernel_fun of infidelity_sampling_MCState
 * (infidelity/overlap/expect.py:78-89, overlap_U/expect.py:101-120):
 *   l_s = T(phi,sigma_s;op1) + T(psi,eta_s;op2) - T(psi,sigma_s) - T(phi,eta_s;op4)
 *   L_s = Re e^{l_s} + cv (e^{2 Re l_s} - 1)          (cv = NaN <-> cv_coeff None)
 * Outputs (device): log_val (S) complex, L (S) real, rho1 (S) complex (weight of the
 * flipped connected element of op2, 0 if op2 is NONE), all of psi->dtype;
 * sums: double[8] ACCUMULATED: [sum L, sum L^2, sum |A|^2, sum Re A, S, n_bad, 0, 0]; n_bad counts the local
 * estimators that came out NaN or infinite (a diverged network: the caller can stop instead of stepping on). */
int nkfb_infidelity_fwd(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi,
                        const nkfb_op_t *op1, const nkfb_op_t *op2, const nkfb_op_t *op4,
                        const uint32_t *sigma, const uint32_t *eta, int64_t S,
                        const double local_states[2], double cv_coeff, void *log_val,
                        void *L, void *rho1, double *sums, void *stream);

/* Parameter tables of the tensor-core forms of nkfb_infidelity_fwd (psi and phi) and nkfb_infidelity_grad (psi), written
 * ahead of time on `stream` -- e.g. under the sampler launches of a step, so that the three small preparation kernels leave
 * the critical path (reference: the parameters are fixed for the whole of one expect_and_grad call,
 * infidelity/overlap_U/expect.py:88-157).  Use with NKFB_OPT_TC_TABLES_VALID = 1; the later launches must be ordered after
 * `stream`.  phi may be NULL (gradient tables only).  Does nothing for shapes / dtypes the tensor-core kernels do not
 * cover. */
int nkfb_tc_prepare(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const double local_states[2],
                    void *stream);

/* ---- infidelity gradient (pass B) --------------------------------------- */
/* Replaces nkjax.expect's backward + nkjax.vjp(conjugate=True)
 * (infidelity/overlap/expect.py:103-125, overlap_U/expect.py:134-156), closed form:
 *   acc_k += sum_s [ w_sig conj O_k(sigma_s) + w_eta conj D_k(eta_s) ],
 *   w_eta = conj A + 2 cv |A|^2,  w_sig = 2 (L_s - F) - w_eta,  F = sums[0]/sums[4]
 * (sums must already hold the GLOBAL sums, i.e. after the scalar all-reduce).
 * grad_acc: device double[2*(M + N*M + N)], complex128 accumulators in the flat layout.
 * op2 of kind NKFB_OP_ISING ((N+1)-connected U^dagger acting on psi): D_k is evaluated on the explicit
 * batch of connected configurations in handle-owned scratch (<= 256 MiB, grown on first use; do not
 * capture that first call in a CUDA graph); rho1 is not read. */
int nkfb_infidelity_grad(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_op_t *op2,
                         const uint32_t *sigma, const uint32_t *eta, int64_t S,
                         const double local_states[2], double cv_coeff,
                         const void *log_val, const void *rho1, const double *sums,
                         double *grad_acc, void *stream);

/* Weighted sum of log-derivatives over an arbitrary list of configurations (the building block of the exact,
 * full-summation infidelity gradient: infidelity/overlap/exact.py:62-76, overlap_U/exact.py:71-88, where
 * nkjax.vjp(conjugate=True) of |<psi|Phi>|^2 / <psi|psi> reduces to such a sum with complex weights):
 *   acc_k += sum_s w_s conj O_k(x_s),   O_k = d log psi / d theta_k.
 * packed: device (S, W32); weights: device (S) complex of net->dtype; acc: device double[2 P], flat layout. */
int nkfb_weighted_score(nkfb_handle_t h, const nkfb_rbm_t *net, const uint32_t *packed, int64_t S,
                        const double local_states[2], const void *weights, double *acc, void *stream);

/* I_grad = -acc / sums[4], cast to `dtype` (infidelity/overlap/expect.py:126-127).
 * grad_out: device (P) complex of `dtype`, flat layout. */
int nkfb_grad_finalize(nkfb_handle_t h, const double *grad_acc, const double *sums,
                       int64_t P, int dtype, void *grad_out, void *stream);

/* One-collective form of the step (reference: ONE mean over ranks, overlap/expect.py:126).
 * acc1 = the pass-B accumulator obtained by handing nkfb_infidelity_grad `centre_sums` = {c0 n, ., ., ., n, ...} in
 * place of the all-reduced sums (any centring constant c0 that is identical on all ranks, e.g. the previous step's F),
 * so that pass B does not have to wait for the global mean; acc2 = sum_s conj O_k(sigma_s) (nkfb_weighted_score with
 * unit weights).  After ONE all-reduce over [sums | acc1 | acc2]:
 * grad_out = -(acc1 + 2 (c0 - F) acc2) / n, F = sums[0] / sums[4] -- exact algebra, not an approximation.
 * All pointers are device pointers. */
int nkfb_grad_finalize_centred(nkfb_handle_t h, const double *acc1, const double *acc2, const double *sums,
                               const double *centre_sums, int64_t P, int dtype, void *grad_out, void *stream);

/* ---- chain statistics ---------------------------------------------------- */
/* Partial sums needed by nk.stats.statistics on data viewed as (rows, cols):
 * per row: [sum, sum of first half, sum of second half] and `n_blocks` block sums of
 * length l_block.  L: device (rows*cols) real of `dtype`.  out: device double
 * [rows * (3 + n_blocks)]. */
int nkfb_chain_stats(nkfb_handle_t h, const void *L, int dtype, int64_t rows, int64_t cols,
                     int64_t l_block, int64_t n_blocks, double *out, void *stream);

/* ---- Renyi-2 swap estimator --------------------------------------------- */
/* Replaces netket.experimental.observable.Renyi2EntanglementEntropy on MCState
 * (netket_fidelity/renyi2.py:1).  samples (2*S_half, W32): first half sigma, second half
 * sigma'.  mask: device (W32) bit i set <-> site i in the partition.
 * q (S_half) complex of net->dtype; sums double[8] accumulated:
 * [sum Re q, sum (Re q)^2, sum Im q, S_half, ...]. */
int nkfb_renyi2(nkfb_handle_t h, const nkfb_rbm_t *net, const uint32_t *samples,
                int64_t S_half, const uint32_t *mask, const double local_states[2],
                uint32_t *scratch /* device (2*S_half, W32) */, void *scratch_logpsi /* device (4*S_half) complex */,
                void *q, double *sums, void *stream);

/* ---- optimiser update ---------------------------------------------------- */
/* Replaces the optax update applied by NetKet's AbstractVariationalDriver.update_parameters
 * (called from InfidelityOptimizer.run, driver/infidelity_optimizer.py:188).  kind 0: SGD,
 * kind 1: Adam (optax.adam; |g|^2 second moment for complex leaves), step t >= 1.
 * params/grad/m: device (P) complex of `dtype` (flat layout); v: device (P) real.
 * real_params != 0: the imaginary part of the gradient is ignored (real-parameter model). */
int nkfb_optimizer_update(nkfb_handle_t h, int kind, int dtype, void *params, const void *grad,
                          void *m, void *v, int64_t P, double lr, double b1, double b2, double eps,
                          int64_t t, int real_params, void *stream);

/* ---- MUFU / FP32 micro-benchmarks (roofline denominators) ---------------- */
/* Jacobian-vector product u_s = sum_k O_k(x_s) v_k of the RBM log-amplitude over a batch (O_k = d log psi / d theta_k;
 * `vec` = P complex numbers of the net's dtype in the flat parameter layout [bias (M) | kernel (N x M) | visible_bias
 * (N)], `out` = S complex numbers).  With nkfb_weighted_score (sum_s w_s conj O_k(x_s)) it gives the action of the
 * quantum geometric tensor S v = <O^* (O v)> - <O^*><O v> without ever forming O: the stochastic-reconfiguration
 * preconditioner behind `preconditioner=` of the reference driver (driver/infidelity_optimizer.py:40,148,205-233;
 * NetKet nk.optimizer.SR / QGTOnTheFly). */
int nkfb_score_dot(nkfb_handle_t h, const nkfb_rbm_t *net, const uint32_t *packed, int64_t S,
                   const double local_states[2], const void *vec, void *out, void *stream);

/* Pipe-rate micro-benchmarks behind the roofline denominators bench.py reports (MEASURED_PEAKS.json holds only the
 * HBM and bf16 tensor peaks): `iters` rounds of 8 independent chains per thread on a full grid of
 * kind 0 MUFU.EX2, 1 MUFU.COS, 2 MUFU.LG2, 3 FFMA with immediate operands, 4 FFMA with three register operands,
 * 5 packed FFMA2 (two FMAs per lane and instruction; counted as FMAs), 6 DFMA (FP64).
 * Returns operations (FMAs) per second measured with CUDA events. */
int nkfb_microbench(nkfb_handle_t h, int kind, int iters, double *ops_per_second, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NKFB200_H */
