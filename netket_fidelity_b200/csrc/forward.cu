// log-amplitude, infidelity estimator (pass A) and gradient (pass B) kernels.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cstdlib>

#include "tile.cuh"

namespace nkfb {

// ------------------------------------------------------------------ shared-memory plan
template <typename R, int SPT>
struct SmemPlan {
  static constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  int N, C;
  __host__ __device__ size_t xs_off() const { return 0; }
  __host__ __device__ size_t ws_off() const { return xs_off() + sizeof(R) * (size_t)(N + 1) * XS; }
  __host__ __device__ size_t amp_off() const { return ws_off() + sizeof(R) * (size_t)(N + 1) * JF; }
  __host__ __device__ size_t t_off() const { return amp_off() + sizeof(R) * 2 * (size_t)TS * C; }
  // T buffers: 5 complex arrays of TS (T1..T4, rho)
  __host__ __device__ size_t total_fwd() const { return t_off() + sizeof(R) * 2 * (size_t)TS * 5; }
};

template <typename R, int SPT>
__device__ __forceinline__ void eval_dispatch(const NetDev<R> &net, const OpDev *op, R *Xs, R *Ws,
                                              cx<R> *amp, cx<R> *Tout, cx<R> *Rout, R s0, R s1) {
  switch (op->kind) {
    case NKFB_OP_GATE:
      eval_tile<R, SPT, NKFB_OP_GATE>(net, op, Xs, Ws, amp, Tout, Rout, s0, s1);
      break;
    case NKFB_OP_ISING:
      eval_tile<R, SPT, NKFB_OP_ISING>(net, op, Xs, Ws, amp, Tout, Rout, s0, s1);
      break;
    default:
      eval_tile<R, SPT, NKFB_OP_NONE>(net, op, Xs, Ws, amp, Tout, Rout, s0, s1);
  }
}

// ------------------------------------------------------------------ logpsi
struct LogpsiArgs {
  OpDev op;
  const uint32_t *packed;
  int64_t B;
  double s0, s1;
  void *out;
  int C;
};

template <typename R, int SPT>
__global__ void __launch_bounds__(NT) logpsi_kernel(NetDev<R> net, LogpsiArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TS = TileCfg<SPT>::TS;
  SmemPlan<R, SPT> pl{net.N, a.C};
  R *Xs = reinterpret_cast<R *>(smem + pl.xs_off());
  R *Ws = reinterpret_cast<R *>(smem + pl.ws_off());
  cx<R> *amp = reinterpret_cast<cx<R> *>(smem + pl.amp_off());
  cx<R> *T = reinterpret_cast<cx<R> *>(smem + pl.t_off());
  __shared__ OpDev op;
  if (threadIdx.x == 0) op = a.op;
  const R s0 = (R)a.s0, s1 = (R)a.s1;
  for (int64_t tile = blockIdx.x; tile * TS < a.B; tile += gridDim.x) {
    const int64_t s_base = tile * TS;
    __syncthreads();
    load_x_tile<R, SPT>(Xs, a.packed, s_base, a.B, net.N, s0, s1);
    eval_dispatch<R, SPT>(net, &op, Xs, Ws, amp, T, nullptr, s0, s1);
    cx<R> *out = reinterpret_cast<cx<R> *>(a.out);
    for (int s = threadIdx.x; s < TS; s += NT)
      if (s_base + s < a.B) out[s_base + s] = T[s];
  }
}

// ------------------------------------------------------------------ pass A
struct FwdArgs {
  OpDev op1, op2, op4;
  const uint32_t *sigma, *eta;
  int64_t S;
  double s0, s1, cv;
  int has_cv;
  void *log_val, *L, *rho1;
  double *sums;
  int C;
};

template <typename R, int SPT>
__global__ void __launch_bounds__(NT) infid_fwd_kernel(NetDev<R> psi, NetDev<R> phi, FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TS = TileCfg<SPT>::TS;
  SmemPlan<R, SPT> pl{psi.N, a.C};
  R *Xs = reinterpret_cast<R *>(smem + pl.xs_off());
  R *Ws = reinterpret_cast<R *>(smem + pl.ws_off());
  cx<R> *amp = reinterpret_cast<cx<R> *>(smem + pl.amp_off());
  cx<R> *T1 = reinterpret_cast<cx<R> *>(smem + pl.t_off());
  cx<R> *T2 = T1 + TS, *T3 = T2 + TS, *T4 = T3 + TS, *RH = T4 + TS;
  __shared__ OpDev ops[4];  // [0]=none [1]=op1 [2]=op2 [3]=op4
  __shared__ double red[8];
  if (threadIdx.x == 0) {
    ops[0].kind = NKFB_OP_NONE;
    ops[1] = a.op1;
    ops[2] = a.op2;
    ops[3] = a.op4;
  }
  const R s0 = (R)a.s0, s1 = (R)a.s1;
  double sL = 0, sL2 = 0, sA2 = 0, sRe = 0, sn = 0, sbad = 0;  // sbad: local estimators that are NaN or infinite
  for (int64_t tile = blockIdx.x; tile * TS < a.S; tile += gridDim.x) {
    const int64_t s_base = tile * TS;
    __syncthreads();
    load_x_tile<R, SPT>(Xs, a.sigma, s_base, a.S, psi.N, s0, s1);
    eval_dispatch<R, SPT>(psi, &ops[0], Xs, Ws, amp, T3, nullptr, s0, s1);
    eval_dispatch<R, SPT>(phi, &ops[1], Xs, Ws, amp, T1, nullptr, s0, s1);
    load_x_tile<R, SPT>(Xs, a.eta, s_base, a.S, psi.N, s0, s1);
    eval_dispatch<R, SPT>(psi, &ops[2], Xs, Ws, amp, T2, RH, s0, s1);
    eval_dispatch<R, SPT>(phi, &ops[3], Xs, Ws, amp, T4, nullptr, s0, s1);
    for (int s = threadIdx.x; s < TS; s += NT) {
      const int64_t gs = s_base + s;
      if (gs < a.S) {
        cx<R> l = T1[s] + T2[s] - T3[s] - T4[s];
        cx<R> A = cexp(l);
        R A2 = abs2(A);
        R Lv = A.re;
        if (a.has_cv) Lv += (R)a.cv * (A2 - R(1));
        reinterpret_cast<cx<R> *>(a.log_val)[gs] = l;
        reinterpret_cast<R *>(a.L)[gs] = Lv;
        reinterpret_cast<cx<R> *>(a.rho1)[gs] = RH[s];
        sL += (double)Lv;
        sL2 += (double)Lv * (double)Lv;
        sA2 += (double)A2;
        sRe += (double)A.re;
        sn += 1.0;
        if (!isfinite((double)Lv)) sbad += 1.0;
      }
    }
  }
  double v[6] = {sL, sL2, sA2, sRe, sn, sbad};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double r = block_sum_256(v[q], red);
    if (threadIdx.x == 0) atomicAdd(a.sums + q, r);
  }
}

// ------------------------------------------------------------------ pass B
struct GradArgs {
  OpDev op2;
  const uint32_t *sigma, *eta;
  int64_t S;
  double s0, s1, cv;
  int has_cv;
  const void *log_val, *rho1;
  const double *sums;
  double *grad_acc;
  int n_splits;
  int64_t tiles_per_split;
  const void *weights;  // nkfb_weighted_score: w_sig = weights[s], no eta term, log_val / rho1 / sums unused
};

template <typename R, int SPT>
struct GradSmem {
  static constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  int N;
  __host__ __device__ size_t xs_off() const { return 0; }
  __host__ __device__ size_t ws_off() const { return sizeof(R) * (size_t)(N + 1) * XS; }
  __host__ __device__ size_t ys_off() const { return ws_off() + sizeof(R) * (size_t)(N + 1) * JF; }
  __host__ __device__ size_t wt_off() const { return ys_off() + sizeof(R) * (size_t)TS * JF; }
  // per-sample weights: w_sig, w_eta, rho1 (complex each)
  __host__ __device__ size_t total() const { return wt_off() + sizeof(R) * 2 * (size_t)TS * 3; }
};

// G[i][f] += sum_s Xs[i][s] * Ys[s][f]; thread (ig = tid/8, jg = tid%8): rows ig + 32 r, cols jg*8..+7
template <typename R, int SPT>
__device__ __forceinline__ void rank1_tile(R G[4][8], const R *Xs, const R *Ys, int nrows) {
  constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  const int ig = threadIdx.x >> 3, jg = threadIdx.x & 7;
  const int nr = (nrows - ig + 31) / 32;  // rows this thread really owns (0..4)
#pragma unroll 1
  for (int s = 0; s < TS; s += 4) {
    R y[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ld4(Ys + (s + q) * JF + jg * 8, y[q]);
      ld4(Ys + (s + q) * JF + jg * 8 + 4, y[q] + 4);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < nr) {
        R x[4];
        ld4(Xs + (ig + 32 * r) * XS + s, x);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int f = 0; f < 8; ++f) G[r][f] = fma(x[q], y[q][f], G[r][f]);
      }
    }
  }
}

template <typename R, int SPT>
__global__ void __launch_bounds__(NT) infid_grad_kernel(NetDev<R> psi, GradArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  GradSmem<R, SPT> pl{psi.N};
  R *Xs = reinterpret_cast<R *>(smem + pl.xs_off());
  R *Ws = reinterpret_cast<R *>(smem + pl.ws_off());
  R *Ys = reinterpret_cast<R *>(smem + pl.ys_off());
  cx<R> *wsig = reinterpret_cast<cx<R> *>(smem + pl.wt_off());
  cx<R> *weta = wsig + TS, *rho = weta + TS;

  const int N = psi.N, M = psi.M;
  const int chunk = blockIdx.x;
  const int rowblk = blockIdx.z;        // rows [128*rowblk, 128*rowblk+128) of the (N+1)-row gradient tile
  const int row0 = rowblk * 128;
  const int nrows = min(128, N + 1 - row0);
  const int ts = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const R s0 = (R)a.s0, s1 = (R)a.s1, ssum = s0 + s1;
  const bool gate = (a.op2.kind == NKFB_OP_GATE);
  const int gidx = a.op2.idx;
  const bool direct = a.weights != nullptr;
  const double F = direct ? 0.0 : a.sums[0] / a.sums[4];
  const R cvc = a.has_cv ? (R)a.cv : R(0);

  load_w_chunk<R>(Ws, psi, chunk);

  R G[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int f = 0; f < 8; ++f) G[r][f] = R(0);
  R corr[4] = {R(0), R(0), R(0), R(0)};

  const int64_t t_begin = (int64_t)blockIdx.y * a.tiles_per_split;
  const int64_t n_tiles = (a.S + TS - 1) / TS;
  const int64_t t_end = min(n_tiles, t_begin + a.tiles_per_split);

  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    const int64_t s_base = tile * TS;
    __syncthreads();  // previous rank-1 done with Xs / Ys / weights
    for (int s = threadIdx.x; s < TS; s += NT) {
      const int64_t gs = s_base + s;
      cx<R> ws_ = mk<R>(0, 0), we_ = mk<R>(0, 0), rh_ = mk<R>(0, 0);
      if (gs < a.S && direct) {
        ws_ = reinterpret_cast<const cx<R> *>(a.weights)[gs];
      } else if (gs < a.S) {
        cx<R> l = reinterpret_cast<const cx<R> *>(a.log_val)[gs];
        cx<R> A = cexp(l);
        R A2 = abs2(A);
        R Lv = A.re + cvc * (A2 - R(1));
        we_ = mk<R>(A.re + R(2) * cvc * A2, -A.im);
        R dl = (R)(2.0 * ((double)Lv - F));
        ws_ = mk<R>(dl - we_.re, -we_.im);
        rh_ = reinterpret_cast<const cx<R> *>(a.rho1)[gs];
      }
      wsig[s] = ws_;
      weta[s] = we_;
      rho[s] = rh_;
    }
    // ---------------- sigma term: Y = w_sig * conj(tanh theta_psi(sigma))
    load_x_tile<R, SPT>(Xs, a.sigma, s_base, a.S, N, s0, s1);
    __syncthreads();
    {
      R acc[SPT][4];
      forward_tile<R, SPT>(acc, Xs, Ws, N);
#pragma unroll
      for (int sp = 0; sp < SPT; ++sp) {
        const int s = ts * SPT + sp;
        const cx<R> w = wsig[s];
        R y[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = chunk * JC + tj * 2 + u;
          cx<R> t = mk<R>(0, 0);
          if (j < M)
            t = conj(ctanh(mk<R>(acc[sp][2 * u], acc[sp][2 * u + 1])));
          else if (j == M)
            t = mk<R>(1, 0);  // pseudo-unit: visible-bias column, O_{a_i} = x_i
          cx<R> v = w * t;
          y[2 * u] = v.re;
          y[2 * u + 1] = v.im;
        }
        st4(Ys + s * JF + tj * 4, y);
      }
    }
    __syncthreads();
    rank1_tile<R, SPT>(G, Xs + row0 * XS, Ys, nrows);
    if (direct) continue;  // uniform over the CTA
    __syncthreads();
    // ---------------- eta term: Y = w_eta * sum_c conj(rho_c) conj(tanh theta_psi(eta'_c))
    load_x_tile<R, SPT>(Xs, a.eta, s_base, a.S, N, s0, s1);
    __syncthreads();
    {
      R acc[SPT][4];
      forward_tile<R, SPT>(acc, Xs, Ws, N);
      R wg[4] = {R(0), R(0), R(0), R(0)};
      if (gate) ld4(Ws + gidx * JF + tj * 4, wg);
#pragma unroll
      for (int sp = 0; sp < SPT; ++sp) {
        const int s = ts * SPT + sp;
        const cx<R> w = weta[s];
        const cx<R> r1 = conj(rho[s]);
        const cx<R> r0 = mk<R>(R(1) - r1.re, -r1.im);
        const R d = gate ? (ssum - R(2) * Xs[gidx * XS + s]) : R(0);
        R y[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = chunk * JC + tj * 2 + u;
          cx<R> v = mk<R>(0, 0);
          if (j <= M) {
            cx<R> t0 = mk<R>(1, 0), t1 = mk<R>(1, 0);
            if (j < M) {
              const cx<R> z = mk<R>(acc[sp][2 * u], acc[sp][2 * u + 1]);
              t0 = conj(ctanh(z));
              if (gate) t1 = conj(ctanh(mk<R>(z.re + d * wg[2 * u], z.im + d * wg[2 * u + 1])));
            }
            if (gate) {
              const cx<R> c1 = w * (r1 * t1);
              v = w * (r0 * t0) + c1;
              corr[2 * u] += d * c1.re;
              corr[2 * u + 1] += d * c1.im;
            } else {
              v = w * t0;
            }
          }
          y[2 * u] = v.re;
          y[2 * u + 1] = v.im;
        }
        st4(Ys + s * JF + tj * 4, y);
      }
    }
    __syncthreads();
    rank1_tile<R, SPT>(G, Xs + row0 * XS, Ys, nrows);
  }

  // ---------------- flush
  const int ig = threadIdx.x >> 3, jg = threadIdx.x & 7;
  const int64_t offK = 2 * (int64_t)M, offA = 2 * ((int64_t)M + (int64_t)N * M);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = row0 + ig + 32 * r;
    if (ig + 32 * r < nrows) {
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        const int col = chunk * JF + jg * 8 + f;
        const int j = col >> 1, comp = col & 1;
        if (j < M) {
          if (i < N)
            atomicAdd(a.grad_acc + offK + 2 * ((int64_t)i * M + j) + comp, (double)G[r][f]);
          else if (psi.b)
            atomicAdd(a.grad_acc + 2 * j + comp, (double)G[r][f]);
        } else if (j == M && i < N && psi.a) {
          atomicAdd(a.grad_acc + offA + 2 * i + comp, (double)G[r][f]);
        }
      }
    }
  }
  if (gate && rowblk == gidx / 128) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int col = chunk * JF + tj * 4 + f;
      const int j = col >> 1, comp = col & 1;
      if (j < M)
        atomicAdd(a.grad_acc + offK + 2 * ((int64_t)gidx * M + j) + comp, (double)corr[f]);
      else if (j == M && psi.a)
        atomicAdd(a.grad_acc + offA + 2 * gidx + comp, (double)corr[f]);
    }
  }
}

template <typename R>
__global__ void grad_finalize_kernel(const double *acc, const double *sums, int64_t P2, R *out) {
  const double sc = -1.0 / sums[4];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P2; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (R)(acc[i] * sc);
}

// ------------------------------------------------------------------ host-side launchers
static int op_conns(const nkfb_op_t *op, int N) {
  if (!op || op->kind == NKFB_OP_NONE) return 1;
  if (op->kind == NKFB_OP_GATE) return 2;
  return N + 1;
}

static int check_net(nkfb_handle_t h, const nkfb_rbm_t *n) {
  if (!n || !n->kernel) NKFB_FAIL(h, NKFB_ERR_INVALID, "null RBM parameters");
  if (n->n_visible < 1 || n->n_hidden < 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad RBM shape");
  if (n->dtype != NKFB_F32 && n->dtype != NKFB_F64) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad RBM dtype");
  return NKFB_OK;
}

static int check_op(nkfb_handle_t h, const nkfb_op_t *op, int N) {
  if (!op) return NKFB_OK;
  if (op->kind < NKFB_OP_NONE || op->kind > NKFB_OP_ISING) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad operator kind");
  if (op->kind == NKFB_OP_GATE && (op->idx < 0 || op->idx >= N))
    NKFB_FAIL(h, NKFB_ERR_INVALID, "gate site %d outside [0,%d)", op->idx, N);
  if (op->kind == NKFB_OP_ISING && op->n_edges > 0 && !op->edges)
    NKFB_FAIL(h, NKFB_ERR_INVALID, "ising operator without edges");
  return NKFB_OK;
}

template <typename K>
static int set_smem(nkfb_handle_t h, K kernel, size_t bytes) {
  if (bytes > 227 * 1024) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "tile needs %zu B of shared memory (> 227 KB)", bytes);
  NKFB_CHECK_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return NKFB_OK;
}

template <typename R, int SPT>
static int launch_logpsi(nkfb_handle_t h, const nkfb_rbm_t *net, const nkfb_op_t *op, const uint32_t *packed,
                         int64_t B, const double ls[2], void *out, cudaStream_t st) {
  constexpr int TS = TileCfg<SPT>::TS;
  LogpsiArgs a;
  a.op = make_opdev(op);
  a.packed = packed;
  a.B = B;
  a.s0 = ls[0];
  a.s1 = ls[1];
  a.out = out;
  a.C = op_conns(op, net->n_visible);
  SmemPlan<R, SPT> pl{net->n_visible, a.C};
  size_t bytes = pl.total_fwd();
  int rc = set_smem(h, logpsi_kernel<R, SPT>, bytes);
  if (rc) return rc;
  int64_t tiles = (B + TS - 1) / TS;
  int grid = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count * 8);
  logpsi_kernel<R, SPT><<<grid, NT, bytes, st>>>(make_netdev<R>(net), a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

template <typename R, int SPT>
static int launch_fwd(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const nkfb_op_t *op1,
                      const nkfb_op_t *op2, const nkfb_op_t *op4, const uint32_t *sigma, const uint32_t *eta,
                      int64_t S, const double ls[2], double cv, void *log_val, void *L, void *rho1, double *sums,
                      cudaStream_t st) {
  constexpr int TS = TileCfg<SPT>::TS;
  FwdArgs a;
  a.op1 = make_opdev(op1);
  a.op2 = make_opdev(op2);
  a.op4 = make_opdev(op4);
  a.sigma = sigma;
  a.eta = eta;
  a.S = S;
  a.s0 = ls[0];
  a.s1 = ls[1];
  a.has_cv = !isnan(cv);
  a.cv = a.has_cv ? cv : 0.0;
  a.log_val = log_val;
  a.L = L;
  a.rho1 = rho1;
  a.sums = sums;
  int N = psi->n_visible;
  a.C = std::max(op_conns(op1, N), std::max(op_conns(op2, N), op_conns(op4, N)));
  SmemPlan<R, SPT> pl{N, a.C};
  size_t bytes = pl.total_fwd();
  int rc = set_smem(h, infid_fwd_kernel<R, SPT>, bytes);
  if (rc) return rc;
  int64_t tiles = (S + TS - 1) / TS;
  int grid = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count * 8);
  infid_fwd_kernel<R, SPT><<<grid, NT, bytes, st>>>(make_netdev<R>(psi), make_netdev<R>(phi), a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

template <typename R, int SPT>
static int launch_grad(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_op_t *op2, const uint32_t *sigma,
                       const uint32_t *eta, int64_t S, const double ls[2], double cv, const void *log_val,
                       const void *rho1, const double *sums, double *grad_acc, cudaStream_t st,
                       const void *weights = nullptr) {
  constexpr int TS = TileCfg<SPT>::TS;
  GradArgs a;
  a.op2 = make_opdev(op2);
  a.sigma = sigma;
  a.eta = eta;
  a.S = S;
  a.s0 = ls[0];
  a.s1 = ls[1];
  a.has_cv = !isnan(cv);
  a.cv = a.has_cv ? cv : 0.0;
  a.log_val = log_val;
  a.rho1 = rho1;
  a.sums = sums;
  a.grad_acc = grad_acc;
  a.weights = weights;
  const int N = psi->n_visible, M = psi->n_hidden;
  const int nchunks = (M + 1 + JC - 1) / JC;  // +1: pseudo-unit column for the visible bias
  const int rowblks = (N + 1 + 127) / 128;
  const int64_t tiles = (S + TS - 1) / TS;
  GradSmem<R, SPT> pl{N};
  size_t bytes = pl.total();
  int rc = set_smem(h, infid_grad_kernel<R, SPT>, bytes);
  if (rc) return rc;
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (bytes + 1024)));
  int64_t want = std::max<int64_t>(1, ((int64_t)h->sm_count * per_sm) / ((int64_t)nchunks * rowblks));
  int64_t n_splits = std::min<int64_t>(tiles, want);
  a.tiles_per_split = (tiles + n_splits - 1) / n_splits;
  n_splits = (tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.n_splits = (int)n_splits;
  dim3 grid(nchunks, (unsigned)n_splits, rowblks);
  infid_grad_kernel<R, SPT><<<grid, NT, bytes, st>>>(make_netdev<R>(psi), a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

// ------------------------------------------------------------------ gradient with an Ising-type U^dagger
// (is_unitary=True with an (N+1)-connected operator acting on psi; reference overlap_U/expect.py:101-120 with
// U_dagger an IsingJax-like DiscreteJaxOperator).  The eta term of the gradient is
//   sum_s w_eta[s] sum_c conj(rho_{s,c}) conj(O_k(eta'_{s,c})),  rho_{s,c} = mel_c psi(eta'_{s,c}) / sum_c' (...)
// and is evaluated on the explicit batch of the (N+1) S connected configurations: bit-packed expansion ->
// logpsi kernel -> per-sample log-sum-exp weights -> the weighted-score form of the gradient kernel.
__global__ void expand_flips_kernel(const uint32_t *eta, int64_t S, int N, uint32_t *out) {
  const int W32 = (N + 31) >> 5, C = N + 1;
  const int64_t total = S * C * W32;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / W32;
    const int w = (int)(idx - row * W32);
    const int64_t s = row / C;
    const int c = (int)(row - s * C);
    uint32_t word = eta[s * W32 + w];
    if (c > 0 && ((c - 1) >> 5) == w) word ^= 1u << ((c - 1) & 31);
    out[idx] = word;
  }
}

// one warp per sample: rho_c from the log-amplitudes of the connected configurations and the matrix elements of
// the operator, then w_sig[s] and w_exp[s, c] = w_eta[s] conj(rho_{s,c})  (same weights as infid_grad_kernel)
template <typename R>
__global__ void ising_weights_kernel(OpDev op, const uint32_t *eta, int64_t S, int N, const cx<R> *lp,
                                     const cx<R> *log_val, const double *sums, R cvc, cx<R> *w_sig,
                                     cx<R> *w_exp) {
  const int W32 = (N + 31) >> 5, C = N + 1;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double F = sums[0] / sums[4];
  for (int64_t s = warp0; s < S; s += nwarps) {
    const cx<R> *am = lp + s * C;
    const uint32_t *x = eta + s * W32;
    R mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmax(mx, am[c].re);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    R zz = R(0);
    for (int e = lane; e < op.n_edges; e += 32) {
      const int ia = op.edges[2 * e], ib = op.edges[2 * e + 1];
      const uint32_t ba = (x[ia >> 5] >> (ia & 31)) & 1u, bb = (x[ib >> 5] >> (ib & 31)) & 1u;
      zz += (ba == bb) ? R(1) : R(-1);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) zz += __shfl_xor_sync(0xffffffffu, zz, m);
    const cx<R> mdiag = mk<R>((R)op.c0[0] + (R)op.c1[0] * zz, (R)op.c0[1] + (R)op.c1[1] * zz);
    const cx<R> moff = mk<R>((R)op.c2[0], (R)op.c2[1]);
    cx<R> sum = mk<R>(0, 0);
    for (int c = lane; c < C; c += 32)
      sum = sum + (c == 0 ? mdiag : moff) * cexp(mk<R>(am[c].re - mx, am[c].im));
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      sum.re += __shfl_xor_sync(0xffffffffu, sum.re, m);
      sum.im += __shfl_xor_sync(0xffffffffu, sum.im, m);
    }
    const cx<R> A = cexp(log_val[s]);
    const R A2 = abs2(A);
    const R Lv = A.re + cvc * (A2 - R(1));
    const cx<R> we = mk<R>(A.re + R(2) * cvc * A2, -A.im);
    if (lane == 0) {
      const R dl = (R)(2.0 * ((double)Lv - F));
      w_sig[s] = mk<R>(dl - we.re, -we.im);
    }
    const R inv = R(1) / abs2(sum);
    for (int c = lane; c < C; c += 32) {
      const cx<R> e = (c == 0 ? mdiag : moff) * cexp(mk<R>(am[c].re - mx, am[c].im));
      const cx<R> rho = inv * (e * conj(sum));
      w_exp[s * C + c] = we * conj(rho);
    }
  }
}

template <typename R, int SPT>
static int launch_grad_ising(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_op_t *op2, const uint32_t *sigma,
                             const uint32_t *eta, int64_t S, const double ls[2], double cv, const void *log_val,
                             const double *sums, double *grad_acc, cudaStream_t st) {
  const int N = psi->n_visible, C = N + 1, W32 = (N + 31) >> 5;
  // scratch per sample: C packed rows, C log-amplitudes, C + 1 weights; chunks of at most ~256 MiB
  const size_t per_sample = (size_t)C * (W32 * 4 + 2 * sizeof(cx<R>)) + sizeof(cx<R>);
  int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(S, (int64_t)((size_t(256) << 20) / per_sample)));
  if (const char *e = getenv("NKFB_ISING_GRAD_CHUNK"))  // testing: force several chunks on small batches
    chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, atoll(e)));
  const size_t need = per_sample * (size_t)chunk + 64;
  if (h->ws4_bytes < need) {
    if (h->ws4) cudaFree(h->ws4);
    h->ws4 = nullptr;
    h->ws4_bytes = 0;
    NKFB_CHECK_CUDA(h, cudaMalloc(&h->ws4, need));
    h->ws4_bytes = need;
  }
  // 16-byte aligned carve-up: complex arrays first
  cx<R> *lp = reinterpret_cast<cx<R> *>(h->ws4);
  cx<R> *w_exp = lp + (size_t)chunk * C;
  cx<R> *w_sig = w_exp + (size_t)chunk * C;
  uint32_t *xp = reinterpret_cast<uint32_t *>(w_sig + chunk);
  const bool has_cv = !isnan(cv);
  const OpDev od = make_opdev(op2);
  for (int64_t s0 = 0; s0 < S; s0 += chunk) {
    const int64_t n = std::min<int64_t>(chunk, S - s0);
    const uint32_t *eta_c = eta + s0 * W32, *sig_c = sigma + s0 * W32;
    const int64_t B = n * C;
    int grid = (int)std::min<int64_t>((B * W32 + 255) / 256, (int64_t)h->sm_count * 16);
    expand_flips_kernel<<<grid, 256, 0, st>>>(eta_c, n, N, xp);
    NKFB_LAUNCHED(h);
    int rc = launch_logpsi<R, SPT>(h, psi, nullptr, xp, B, ls, lp, st);
    if (rc) return rc;
    grid = (int)std::min<int64_t>((n * 32 + 255) / 256, (int64_t)h->sm_count * 16);
    ising_weights_kernel<R><<<grid, 256, 0, st>>>(od, eta_c, n, N, lp,
                                                  reinterpret_cast<const cx<R> *>(log_val) + s0, sums,
                                                  has_cv ? (R)cv : R(0), w_sig, w_exp);
    NKFB_LAUNCHED(h);
    rc = launch_grad<R, SPT>(h, psi, nullptr, sig_c, sig_c, n, ls, NAN, nullptr, nullptr, nullptr, grad_acc, st,
                             w_sig);
    if (rc) return rc;
    rc = launch_grad<R, SPT>(h, psi, nullptr, xp, xp, B, ls, NAN, nullptr, nullptr, nullptr, grad_acc, st, w_exp);
    if (rc) return rc;
  }
  return NKFB_OK;
}

}  // namespace nkfb

using namespace nkfb;

extern "C" int nkfb_logpsi(nkfb_handle_t h, const nkfb_rbm_t *net, const nkfb_op_t *op, const uint32_t *packed,
                           int64_t B, const double local_states[2], void *out, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_logpsi");
  int rc = check_net(h, net);
  if (rc) return rc;
  rc = check_op(h, op, net->n_visible);
  if (rc) return rc;
  if (B <= 0) return NKFB_OK;
  if (!packed || !out) NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool ising = op && op->kind == NKFB_OP_ISING;
  if (net->dtype == NKFB_F32)
    return ising ? launch_logpsi<float, 4>(h, net, op, packed, B, local_states, out, st)
                 : launch_logpsi<float, 8>(h, net, op, packed, B, local_states, out, st);
  return launch_logpsi<double, 4>(h, net, op, packed, B, local_states, out, st);
}

namespace nkfb {
int tc_infidelity_fwd(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const nkfb_op_t *op1,
                      const nkfb_op_t *op2, const nkfb_op_t *op4, const uint32_t *sigma, const uint32_t *eta,
                      int64_t S, const double ls[2], double cv, void *log_val, void *L, void *rho1, double *sums,
                      cudaStream_t st);
int tc_infidelity_grad(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_op_t *op2, const uint32_t *sigma,
                       const uint32_t *eta, int64_t S, const double ls[2], double cv, const void *log_val,
                       const void *rho1, const double *sums, double *grad_acc, cudaStream_t st,
                       const void *weights = nullptr);
int tc_fwd_prepare(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const double ls[2], cudaStream_t st);
int tc_grad_prepare(nkfb_handle_t h, const nkfb_rbm_t *psi, const double ls[2], cudaStream_t st);
}

extern "C" int nkfb_tc_prepare(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi,
                               const double local_states[2], void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_tc_prepare");
  int rc = check_net(h, psi);
  if (rc) return rc;
  if (phi && (rc = check_net(h, phi))) return rc;
  if (!local_states) NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (phi && psi->n_visible == phi->n_visible && psi->dtype == phi->dtype) {
    const char *impl = getenv("NKFB_FWD_IMPL");
    if (!(impl && strcmp(impl, "simt") == 0) && (rc = nkfb::tc_fwd_prepare(h, psi, phi, local_states, st))) return rc;
  }
  const char *gimpl = getenv("NKFB_GRAD_IMPL");
  if (!(gimpl && strcmp(gimpl, "simt") == 0) && (rc = nkfb::tc_grad_prepare(h, psi, local_states, st))) return rc;
  return NKFB_OK;
}

extern "C" int nkfb_infidelity_fwd(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi,
                                   const nkfb_op_t *op1, const nkfb_op_t *op2, const nkfb_op_t *op4,
                                   const uint32_t *sigma, const uint32_t *eta, int64_t S,
                                   const double local_states[2], double cv_coeff, void *log_val, void *L,
                                   void *rho1, double *sums, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_infidelity_fwd");
  int rc = check_net(h, psi);
  if (rc) return rc;
  rc = check_net(h, phi);
  if (rc) return rc;
  if (psi->n_visible != phi->n_visible) NKFB_FAIL(h, NKFB_ERR_INVALID, "Hilbert spaces should match");
  if (psi->dtype != phi->dtype) NKFB_FAIL(h, NKFB_ERR_INVALID, "psi and phi must share a dtype");
  const int N = psi->n_visible;
  if ((rc = check_op(h, op1, N)) || (rc = check_op(h, op2, N)) || (rc = check_op(h, op4, N))) return rc;
  if (S <= 0) return NKFB_OK;
  if (!sigma || !eta || !log_val || !L || !rho1 || !sums) NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool ising = (op1 && op1->kind == NKFB_OP_ISING) || (op2 && op2->kind == NKFB_OP_ISING) ||
                     (op4 && op4->kind == NKFB_OP_ISING);
  {
    // tensor-core path (fp32, NONE/GATE operators, large batches); NKFB_FWD_IMPL=simt forces the CUDA-core kernel
    const char *impl = getenv("NKFB_FWD_IMPL");
    if (!(impl && strcmp(impl, "simt") == 0)) {
      rc = tc_infidelity_fwd(h, psi, phi, op1, op2, op4, sigma, eta, S, local_states, cv_coeff, log_val, L, rho1,
                             sums, st);
      if (rc <= 0) return rc;  // done (0) or error (< 0); 1 = not applicable
    }
  }
  if (psi->dtype == NKFB_F32)
    return ising ? launch_fwd<float, 4>(h, psi, phi, op1, op2, op4, sigma, eta, S, local_states, cv_coeff, log_val,
                                        L, rho1, sums, st)
                 : launch_fwd<float, 8>(h, psi, phi, op1, op2, op4, sigma, eta, S, local_states, cv_coeff, log_val,
                                        L, rho1, sums, st);
  return launch_fwd<double, 4>(h, psi, phi, op1, op2, op4, sigma, eta, S, local_states, cv_coeff, log_val, L, rho1,
                               sums, st);
}

extern "C" int nkfb_infidelity_grad(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_op_t *op2,
                                    const uint32_t *sigma, const uint32_t *eta, int64_t S,
                                    const double local_states[2], double cv_coeff, const void *log_val,
                                    const void *rho1, const double *sums, double *grad_acc, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_infidelity_grad");
  int rc = check_net(h, psi);
  if (rc) return rc;
  rc = check_op(h, op2, psi->n_visible);
  if (rc) return rc;
  if (S <= 0) return NKFB_OK;
  if (!sigma || !eta || !log_val || !rho1 || !sums || !grad_acc) NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (op2 && op2->kind == NKFB_OP_ISING) {
    // (N+1)-connected U^dagger acting on psi: explicit connected batch (rho1 is not used)
    if (psi->dtype == NKFB_F32)
      return launch_grad_ising<float, 8>(h, psi, op2, sigma, eta, S, local_states, cv_coeff, log_val, sums, grad_acc,
                                         st);
    return launch_grad_ising<double, 4>(h, psi, op2, sigma, eta, S, local_states, cv_coeff, log_val, sums, grad_acc,
                                        st);
  }
  {
    const char *impl = getenv("NKFB_GRAD_IMPL");
    if (!(impl && strcmp(impl, "simt") == 0)) {
      rc = tc_infidelity_grad(h, psi, op2, sigma, eta, S, local_states, cv_coeff, log_val, rho1, sums, grad_acc, st);
      if (rc <= 0) return rc;
    }
  }
  if (psi->dtype == NKFB_F32)
    return launch_grad<float, 8>(h, psi, op2, sigma, eta, S, local_states, cv_coeff, log_val, rho1, sums, grad_acc,
                                 st);
  return launch_grad<double, 4>(h, psi, op2, sigma, eta, S, local_states, cv_coeff, log_val, rho1, sums, grad_acc,
                                st);
}

extern "C" int nkfb_weighted_score(nkfb_handle_t h, const nkfb_rbm_t *net, const uint32_t *packed, int64_t S,
                                   const double local_states[2], const void *weights, double *acc, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_weighted_score");
  int rc = check_net(h, net);
  if (rc) return rc;
  if (S <= 0) return NKFB_OK;
  if (!packed || !weights || !acc) NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // tensor-core path (single precision, S >= 2048): the sigma side of pass B with the given weights
    const char *impl = getenv("NKFB_GRAD_IMPL");
    if (!(impl && strcmp(impl, "simt") == 0)) {
      rc = tc_infidelity_grad(h, net, nullptr, packed, packed, S, local_states, NAN, nullptr, nullptr, nullptr, acc, st,
                              weights);
      if (rc <= 0) return rc;
    }
  }
  if (net->dtype == NKFB_F32)
    return launch_grad<float, 8>(h, net, nullptr, packed, packed, S, local_states, NAN, nullptr, nullptr, nullptr, acc,
                                 st, weights);
  return launch_grad<double, 4>(h, net, nullptr, packed, packed, S, local_states, NAN, nullptr, nullptr, nullptr, acc,
                                st, weights);
}

namespace nkfb {
// one-collective form: acc1 was accumulated with the centring constant c0 instead of the global mean F,
// acc2 = sum_s conj O_k(sigma_s); I_grad = -(acc1 + 2 (c0 - F) acc2) / n with F = sums[0] / sums[4] AFTER the all-reduce
template <typename R>
__global__ void grad_finalize_centred_kernel(const double *acc1, const double *acc2, const double *sums,
                                             const double *centre, int64_t n2, R *out) {
  const double n = sums[4], F = sums[0] / n, c0 = centre[0] / centre[4];
  const double k = 2.0 * (c0 - F);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (R)(-(acc1[i] + k * acc2[i]) / n);
}
}  // namespace nkfb

extern "C" int nkfb_grad_finalize_centred(nkfb_handle_t h, const double *acc1, const double *acc2, const double *sums,
                                          const double *centre_sums, int64_t P, int dtype, void *grad_out,
                                          void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_grad_finalize_centred");
  if (!acc1 || !acc2 || !sums || !centre_sums || !grad_out || P <= 0) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad finalize arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)std::min<int64_t>((2 * P + 255) / 256, 1024);
  if (dtype == NKFB_F32)
    grad_finalize_centred_kernel<float><<<grid, 256, 0, st>>>(acc1, acc2, sums, centre_sums, 2 * P, (float *)grad_out);
  else if (dtype == NKFB_F64)
    grad_finalize_centred_kernel<double><<<grid, 256, 0, st>>>(acc1, acc2, sums, centre_sums, 2 * P, (double *)grad_out);
  else
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad dtype");
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

extern "C" int nkfb_grad_finalize(nkfb_handle_t h, const double *grad_acc, const double *sums, int64_t P, int dtype,
                                  void *grad_out, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_grad_finalize");
  if (!grad_acc || !sums || !grad_out || P <= 0) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad finalize arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)std::min<int64_t>((2 * P + 255) / 256, 1024);
  if (dtype == NKFB_F32)
    grad_finalize_kernel<float><<<grid, 256, 0, st>>>(grad_acc, sums, 2 * P, (float *)grad_out);
  else if (dtype == NKFB_F64)
    grad_finalize_kernel<double><<<grid, 256, 0, st>>>(grad_acc, sums, 2 * P, (double *)grad_out);
  else
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad dtype");
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}
