// Jacobian-vector product of the RBM log-amplitude over a batch of configurations:
//     u_s = sum_k O_k(x_s) v_k,     O_k(x) = d log psi(x) / d theta_k
// the half of the quantum-geometric-tensor product  S v = <O^* (O v)> - <O^*><O v>  that pass B's machinery does
// not have (the other half, sum_s w_s conj O_k(x_s), is nkfb_weighted_score).  Used by the stochastic-reconfiguration
// preconditioner (netket_fidelity_b200/nk/optimizer.py: SR; reference hook: `preconditioner=` of
// netket_fidelity/driver/infidelity_optimizer.py:40,148,205-233, NetKet nk.optimizer.SR with QGTOnTheFly).
//
// For the RBM   u_s = sum_j tanh(theta_sj) (vb_j + sum_i x_si vW_ij) + sum_i x_si va_i,   theta = b + x W,
// i.e. two forward passes over the same X tile -- one with the parameters, one with the vector laid out like the
// parameters (flat layout [bias | kernel | visible_bias]) -- combined unit by unit: the tile machinery of tile.cuh.
#include <cstdio>
#include <cstring>
#include <algorithm>

#include "tile.cuh"

namespace nkfb {

template <typename R, int SPT>
__global__ void __launch_bounds__(NT) score_dot_kernel(NetDev<R> net, NetDev<R> vec, const uint32_t *packed, int64_t S,
                                                        R s0, R s1, cx<R> *out) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  const int N = net.N, M = net.M;
  R *Xs = reinterpret_cast<R *>(smem);
  R *Ws = Xs + (size_t)(N + 1) * XS;
  const int ts = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int nchunks = (M + JC - 1) / JC;
  for (int64_t tile = blockIdx.x; tile * TS < S; tile += gridDim.x) {
    const int64_t s_base = tile * TS;
    __syncthreads();
    load_x_tile<R, SPT>(Xs, packed, s_base, S, N, s0, s1);
    cx<R> u[SPT];
#pragma unroll
    for (int sp = 0; sp < SPT; ++sp) u[sp] = mk<R>(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      __syncthreads();
      load_w_chunk<R>(Ws, net, chunk);
      __syncthreads();
      R acc[SPT][4];
      forward_tile<R, SPT>(acc, Xs, Ws, N);
      cx<R> t[SPT][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const bool live = chunk * JC + tj * 2 + q < M;
#pragma unroll
        for (int sp = 0; sp < SPT; ++sp)
          t[sp][q] = live ? ctanh(mk<R>(acc[sp][2 * q], acc[sp][2 * q + 1])) : mk<R>(0, 0);
      }
      __syncthreads();
      load_w_chunk<R>(Ws, vec, chunk);   // row N = the bias part of the vector
      __syncthreads();
      forward_tile<R, SPT>(acc, Xs, Ws, N);
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int sp = 0; sp < SPT; ++sp) u[sp] = u[sp] + t[sp][q] * mk<R>(acc[sp][2 * q], acc[sp][2 * q + 1]);
    }
    if (vec.a) {
      for (int i = tj; i < N; i += 16) {
        const cx<R> ai = mk<R>(vec.a[2 * i], vec.a[2 * i + 1]);
#pragma unroll
        for (int sp = 0; sp < SPT; ++sp) u[sp] = u[sp] + Xs[i * XS + ts * SPT + sp] * ai;
      }
    }
#pragma unroll
    for (int sp = 0; sp < SPT; ++sp) {
      u[sp].re = half_warp_sum(u[sp].re);
      u[sp].im = half_warp_sum(u[sp].im);
      const int64_t gs = s_base + ts * SPT + sp;
      if (tj == 0 && gs < S) out[gs] = u[sp];
    }
  }
}

template <typename R, int SPT>
static int launch_score_dot(nkfb_handle_t h, const nkfb_rbm_t *net, const void *vec, const uint32_t *packed, int64_t S,
                            const double ls[2], void *out, cudaStream_t st) {
  constexpr int TS = TileCfg<SPT>::TS, XS = TileCfg<SPT>::XS;
  const int N = net->n_visible, M = net->n_hidden;
  NetDev<R> nd = make_netdev<R>(net), vd;
  const R *v = (const R *)vec;  // flat layout [bias (M) | kernel (N x M) | visible_bias (N)], complex interleaved
  vd.N = N;
  vd.M = M;
  vd.b = net->bias ? v : nullptr;
  vd.W = v + 2 * (size_t)M;
  vd.a = net->visible_bias ? v + 2 * ((size_t)M + (size_t)N * M) : nullptr;
  const size_t bytes = sizeof(R) * ((size_t)(N + 1) * XS + (size_t)(N + 1) * JF);
  if (bytes > 227 * 1024) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "tile needs %zu B of shared memory (> 227 KB)", bytes);
  NKFB_CHECK_CUDA(h, cudaFuncSetAttribute(score_dot_kernel<R, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const int64_t tiles = (S + TS - 1) / TS;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count * 8);
  score_dot_kernel<R, SPT><<<grid, NT, bytes, st>>>(nd, vd, packed, S, (R)ls[0], (R)ls[1], (cx<R> *)out);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb

using namespace nkfb;

extern "C" int nkfb_score_dot(nkfb_handle_t h, const nkfb_rbm_t *net, const uint32_t *packed, int64_t S,
                              const double local_states[2], const void *vec, void *out, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_score_dot");
  if (!net || !net->kernel || net->n_visible < 1 || net->n_hidden < 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad RBM");
  if (net->dtype != NKFB_F32 && net->dtype != NKFB_F64) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad RBM dtype");
  if (S <= 0) return NKFB_OK;
  if (!packed || !vec || !out) NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (net->dtype == NKFB_F32) return launch_score_dot<float, 8>(h, net, vec, packed, S, local_states, out, st);
  return launch_score_dot<double, 4>(h, net, vec, packed, S, local_states, out, st);
}
