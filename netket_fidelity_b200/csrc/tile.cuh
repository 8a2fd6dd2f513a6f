// Sample-tile x hidden-chunk machinery shared by the log-amplitude, estimator (pass A)
// and gradient (pass B) kernels.
//
// A CTA of 256 threads works on a tile of TS = 16*SPT configurations and one chunk of
// JC = 32 hidden units (JF = 64 reals, complex interleaved) at a time:
//   Theta[s, j] = b_j + sum_i x_{s,i} W_{i,j}
// is a (TS x (N+1)) * ((N+1) x JF) real product with the bias folded in as row N
// (x_{s,N} = 1).  Thread (ts = tid/16, tj = tid%16) owns SPT samples x 4 reals
// (2 complex hidden units) of the result.
#pragma once

#include "common.cuh"

namespace nkfb {

constexpr int NT = 256;
constexpr int JF = 64;  // reals per hidden chunk
constexpr int JC = 32;  // complex hidden units per chunk

template <typename R>
struct NetDev {
  int N, M;
  const R *W;  // (N, M) complex interleaved
  const R *b;  // (M) or null
  const R *a;  // (N) or null
};

struct OpDev {
  int kind, idx, n_edges;
  const int32_t *edges;
  double gate_mel[2][2][2];
  double c0[2], c1[2], c2[2];
};

inline OpDev make_opdev(const nkfb_op_t *op) {
  OpDev o;
  memset(&o, 0, sizeof(o));
  if (op == nullptr) {
    o.kind = NKFB_OP_NONE;
    return o;
  }
  o.kind = op->kind;
  o.idx = op->idx;
  o.n_edges = op->n_edges;
  o.edges = op->edges;
  memcpy(o.gate_mel, op->gate_mel, sizeof(o.gate_mel));
  memcpy(o.c0, op->c0, sizeof(o.c0));
  memcpy(o.c1, op->c1, sizeof(o.c1));
  memcpy(o.c2, op->c2, sizeof(o.c2));
  return o;
}

template <typename R>
inline NetDev<R> make_netdev(const nkfb_rbm_t *n) {
  NetDev<R> d;
  d.N = n->n_visible;
  d.M = n->n_hidden;
  d.W = (const R *)n->kernel;
  d.b = (const R *)n->bias;
  d.a = (const R *)n->visible_bias;
  return d;
}

template <int SPT>
struct TileCfg {
  static constexpr int TS = 16 * SPT;  // samples per tile
  static constexpr int XS = TS + 4;    // padded row stride of the X tile (keeps 16B alignment,
                                       // de-conflicts the rank-1 phase of pass B)
};

__device__ __forceinline__ void ld4(const float *p, float o[4]) {
  float4 v = *reinterpret_cast<const float4 *>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void ld4(const double *p, double o[4]) {
  double2 a = *reinterpret_cast<const double2 *>(p);
  double2 b = *reinterpret_cast<const double2 *>(p + 2);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void st4(float *p, const float o[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void st4(double *p, const double o[4]) {
  *reinterpret_cast<double2 *>(p) = make_double2(o[0], o[1]);
  *reinterpret_cast<double2 *>(p + 2) = make_double2(o[2], o[3]);
}

// X tile: Xs[i * XS + s] = local-state value of site i of sample s_base + s (0 past the end),
// row N = 1 (bias row).
template <typename R, int SPT>
__device__ __forceinline__ void load_x_tile(R *Xs, const uint32_t *packed, int64_t s_base, int64_t S,
                                            int N, R s0, R s1) {
  constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  const int W32 = (N + 31) >> 5;
  for (int idx = threadIdx.x; idx < (N + 1) * TS; idx += NT) {
    int i = idx / TS, s = idx - i * TS;
    int64_t gs = s_base + s;
    R v = R(0);
    if (i == N) {
      v = R(1);
    } else if (gs < S) {
      uint32_t w = packed[gs * W32 + (i >> 5)];
      v = ((w >> (i & 31)) & 1u) ? s1 : s0;
    }
    Xs[i * XS + s] = v;
  }
}

// W chunk: Ws[i * JF + f] = real column (chunk*JF + f) of row i of the kernel; row N = bias.
template <typename R>
__device__ __forceinline__ void load_w_chunk(R *Ws, const NetDev<R> &net, int chunk) {
  const int N = net.N, M2 = 2 * net.M;
  for (int idx = threadIdx.x; idx < (N + 1) * JF; idx += NT) {
    int i = idx / JF, f = idx - i * JF;
    int col = chunk * JF + f;
    R v = R(0);
    if (col < M2) {
      if (i < N)
        v = net.W[(int64_t)i * M2 + col];
      else if (net.b)
        v = net.b[col];
    }
    Ws[idx] = v;
  }
}

// acc[sp][f] = sum_{i<=N} Xs[i][ts*SPT+sp] * Ws[i][tj*4+f]
template <typename R, int SPT>
__device__ __forceinline__ void forward_tile(R acc[SPT][4], const R *Xs, const R *Ws, int N) {
  constexpr int XS = TileCfg<SPT>::XS;
  const int ts = threadIdx.x >> 4, tj = threadIdx.x & 15;
#pragma unroll
  for (int sp = 0; sp < SPT; ++sp)
#pragma unroll
    for (int f = 0; f < 4; ++f) acc[sp][f] = R(0);
  const R *xp = Xs + ts * SPT;
  const R *wp = Ws + tj * 4;
#pragma unroll 2
  for (int i = 0; i <= N; ++i) {
    R xv[SPT], wv[4];
#pragma unroll
    for (int q = 0; q < SPT; q += 4) ld4(xp + i * XS + q, xv + q);
    ld4(wp + i * JF, wv);
#pragma unroll
    for (int sp = 0; sp < SPT; ++sp)
#pragma unroll
      for (int f = 0; f < 4; ++f) acc[sp][f] = fma(xv[sp], wv[f], acc[sp][f]);
  }
}

// sum over the 16 lanes that share a sample group (lanes 0-15 / 16-31 of a warp)
template <typename R>
__device__ __forceinline__ R half_warp_sum(R v) {
#pragma unroll
  for (int m = 8; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ---------------------------------------------------------------------------------------
// eval_tile: for the X tile currently in Xs, computes for every sample s of the tile
//   T[s]    = log sum_c mel_c(x_s) * net(x'_{s,c})          (complex)
//   rho1[s] = mel_1 net(x'_{s,1}) / sum_c ...               (GATE only; 0 otherwise)
// into shared arrays Tout / Rout (may be null).  amp: scratch TS*(C) complex in smem
// (C = 1, 2 or N+1).  All threads of the CTA must call it.
// ---------------------------------------------------------------------------------------
template <typename R, int SPT, int OPK>
__device__ __noinline__ void eval_tile(const NetDev<R> net, const OpDev *op, R *Xs, R *Ws,
                                       cx<R> *amp, cx<R> *Tout, cx<R> *Rout, R s0, R s1) {
  constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  const int N = net.N, M = net.M;
  const int ts = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int nchunks = (M + JC - 1) / JC;
  const int C = (OPK == NKFB_OP_NONE) ? 1 : (OPK == NKFB_OP_GATE ? 2 : N + 1);
  const R ssum = s0 + s1;

  cx<R> a0[SPT], a1[SPT];
  R dfl[SPT];
#pragma unroll
  for (int sp = 0; sp < SPT; ++sp) {
    a0[sp] = mk<R>(0, 0);
    a1[sp] = mk<R>(0, 0);
    dfl[sp] = R(0);
  }
  if (OPK == NKFB_OP_ISING) {
    for (int idx = threadIdx.x; idx < TS * C; idx += NT) amp[idx] = mk<R>(0, 0);
  }
  __syncthreads();  // X tile (and amp zeroing) visible
  if (OPK == NKFB_OP_GATE) {
#pragma unroll
    for (int sp = 0; sp < SPT; ++sp) dfl[sp] = ssum - R(2) * Xs[op->idx * XS + ts * SPT + sp];
  }

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    __syncthreads();
    load_w_chunk<R>(Ws, net, chunk);
    __syncthreads();
    R acc[SPT][4];
    forward_tile<R, SPT>(acc, Xs, Ws, N);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = chunk * JC + tj * 2 + u;
      if (j < M) {
        cx<R> wg = mk<R>(0, 0);
        if (OPK == NKFB_OP_GATE) wg = mk<R>(Ws[op->idx * JF + tj * 4 + 2 * u], Ws[op->idx * JF + tj * 4 + 2 * u + 1]);
#pragma unroll
        for (int sp = 0; sp < SPT; ++sp) {
          cx<R> z = mk<R>(acc[sp][2 * u], acc[sp][2 * u + 1]);
          a0[sp] = a0[sp] + lncosh(z);
          if (OPK == NKFB_OP_GATE) a1[sp] = a1[sp] + lncosh(z + dfl[sp] * wg);
        }
      }
    }
    if (OPK == NKFB_OP_ISING) {
      // every single-flip neighbour k: theta' = theta + d_{s,k} W[k, :]
      for (int k = 0; k < N; ++k) {
        R wk[4];
        ld4(Ws + k * JF + tj * 4, wk);
#pragma unroll
        for (int sp = 0; sp < SPT; ++sp) {
          R d = ssum - R(2) * Xs[k * XS + ts * SPT + sp];
          cx<R> part = mk<R>(0, 0);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int j = chunk * JC + tj * 2 + u;
            if (j < M) {
              cx<R> z = mk<R>(acc[sp][2 * u] + d * wk[2 * u], acc[sp][2 * u + 1] + d * wk[2 * u + 1]);
              part = part + lncosh(z);
            }
          }
          part.re = half_warp_sum(part.re);
          part.im = half_warp_sum(part.im);
          if (tj == 0) {
            cx<R> *dst = amp + (ts * SPT + sp) * C + (k + 1);
            *dst = *dst + part;
          }
        }
      }
    }
  }

  // visible-bias term a.x, distributed over the 16 lanes of the sample group
  cx<R> ax[SPT];
#pragma unroll
  for (int sp = 0; sp < SPT; ++sp) ax[sp] = mk<R>(0, 0);
  if (net.a) {
    for (int i = tj; i < N; i += 16) {
      cx<R> ai = mk<R>(net.a[2 * i], net.a[2 * i + 1]);
#pragma unroll
      for (int sp = 0; sp < SPT; ++sp) ax[sp] = ax[sp] + Xs[i * XS + ts * SPT + sp] * ai;
    }
  }
#pragma unroll
  for (int sp = 0; sp < SPT; ++sp) {
    a0[sp] = a0[sp] + ax[sp];
    a0[sp].re = half_warp_sum(a0[sp].re);
    a0[sp].im = half_warp_sum(a0[sp].im);
    if (OPK != NKFB_OP_NONE) {
      ax[sp].re = half_warp_sum(ax[sp].re);
      ax[sp].im = half_warp_sum(ax[sp].im);
    }
    if (OPK == NKFB_OP_GATE) {
      a1[sp].re = half_warp_sum(a1[sp].re);
      a1[sp].im = half_warp_sum(a1[sp].im);
    }
  }
  // a0 = log net(x) ; a1 + ax + d a_idx = log net(x flipped at idx)

  if (OPK == NKFB_OP_NONE) {
    if (tj == 0) {
#pragma unroll
      for (int sp = 0; sp < SPT; ++sp) {
        Tout[ts * SPT + sp] = a0[sp];
        if (Rout) Rout[ts * SPT + sp] = mk<R>(0, 0);
      }
    }
  } else if (OPK == NKFB_OP_GATE) {
    if (tj == 0) {
      cx<R> aidx = net.a ? mk<R>(net.a[2 * op->idx], net.a[2 * op->idx + 1]) : mk<R>(0, 0);
#pragma unroll
      for (int sp = 0; sp < SPT; ++sp) {
        const int s = ts * SPT + sp;
        const R xi = Xs[op->idx * XS + s];
        const int bit = (xi == s1) ? 1 : 0;
        const cx<R> m0 = mk<R>((R)op->gate_mel[bit][0][0], (R)op->gate_mel[bit][0][1]);
        const cx<R> m1 = mk<R>((R)op->gate_mel[bit][1][0], (R)op->gate_mel[bit][1][1]);
        const cx<R> l0 = a0[sp];
        const cx<R> l1 = a1[sp] + ax[sp] + dfl[sp] * aidx;
        // logsumexp(a, b=mels): shift by max Re a, complex sum, principal log
        const R amax = fmax(l0.re, l1.re);
        const cx<R> e0 = m0 * cexp(mk<R>(l0.re - amax, l0.im));
        const cx<R> e1 = m1 * cexp(mk<R>(l1.re - amax, l1.im));
        const cx<R> sum = e0 + e1;
        cx<R> T = clog(sum);
        T.re += amax;
        Tout[s] = T;
        if (Rout) {
          const R inv = R(1) / abs2(sum);
          Rout[s] = inv * (e1 * conj(sum));
        }
      }
    }
  } else {  // ISING
    if (tj == 0) {
#pragma unroll
      for (int sp = 0; sp < SPT; ++sp) amp[(ts * SPT + sp) * C] = a0[sp];
    }
    __syncwarp();
    // finish the neighbour amplitudes and do the log-sum-exp, 16 lanes per sample group
#pragma unroll
    for (int sp = 0; sp < SPT; ++sp) {
      const int s = ts * SPT + sp;
      cx<R> *am = amp + s * C;
      R mx = -INFINITY;
      for (int c = tj; c < C; c += 16) {
        if (c > 0) {
          const int k = c - 1;
          const R d = ssum - R(2) * Xs[k * XS + s];
          cx<R> v = am[c] + ax[sp];
          if (net.a) v = v + d * mk<R>(net.a[2 * k], net.a[2 * k + 1]);
          am[c] = v;
        }
        mx = fmax(mx, am[c].re);
      }
#pragma unroll
      for (int m = 8; m > 0; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
      // diagonal matrix element: c0 + c1 * sum_<ij> (2 [x_i == x_j] - 1)
      R zz = R(0);
      for (int e = tj; e < op->n_edges; e += 16) {
        const R xa = Xs[op->edges[2 * e] * XS + s], xb = Xs[op->edges[2 * e + 1] * XS + s];
        zz += (xa == xb) ? R(1) : R(-1);
      }
      zz = half_warp_sum(zz);
      const cx<R> mdiag = mk<R>((R)op->c0[0] + (R)op->c1[0] * zz, (R)op->c0[1] + (R)op->c1[1] * zz);
      const cx<R> moff = mk<R>((R)op->c2[0], (R)op->c2[1]);
      cx<R> sum = mk<R>(0, 0);
      for (int c = tj; c < C; c += 16) {
        const cx<R> e = cexp(mk<R>(am[c].re - mx, am[c].im));
        sum = sum + (c == 0 ? mdiag : moff) * e;
      }
      sum.re = half_warp_sum(sum.re);
      sum.im = half_warp_sum(sum.im);
      if (tj == 0) {
        cx<R> T = clog(sum);
        T.re += mx;
        Tout[s] = T;
        if (Rout) Rout[s] = mk<R>(0, 0);
      }
    }
  }
  __syncthreads();
}

}  // namespace nkfb
