// Metropolis-local sampler for a plain RBM target |psi|^2 (see sampler.cu for the overview).
//
// Template parameters: R real type; L lanes per chain group (4/8/16/32); K hidden units per
// lane, EXACT (K = ceil(M / L): no per-unit predicates in the Metropolis loop, only the last
// unit of a lane can fall outside M and it is handled by one predicated load of W);
// C chains per lane group (they share the row W[i, :] when the site sequence is shared).
#pragma once

#include "sampler_common.cuh"

namespace nkfb {

template <typename R>
struct Vec2;
template <>
struct Vec2<float> {
  using type = float2;
};
template <>
struct Vec2<double> {
  using type = double2;
};

template <typename R, int L, int K>
struct RowLoader {
  using V2 = typename Vec2<R>::type;
  const V2 *W;  // complex rows, M per row
  int M, l;
  bool last_ok;
  __device__ __forceinline__ RowLoader(const R *Wp, int M_, int l_) : W(reinterpret_cast<const V2 *>(Wp)), M(M_), l(l_) {
    last_ok = (l + L * (K - 1)) < M;
  }
  __device__ __forceinline__ void load(int row, R (&wr)[K], R (&wi)[K]) const {
    const V2 *p = W + (int64_t)row * M + l;
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
      const V2 v = __ldg(p + L * k);
      wr[k] = v.x;
      wi[k] = v.y;
    }
    V2 v;
    v.x = R(0);
    v.y = R(0);
    if (last_ok) v = __ldg(p + L * (K - 1));
    wr[K - 1] = v.x;
    wi[K - 1] = v.y;
  }
};

// sum_j log2|cosh(theta_j)| + (number of units), in the log2 domain, for the K units of this lane
template <typename R, int K>
__device__ __forceinline__ R lane_logcosh2(const R (&xr)[K], const R (&xi)[K]) {
  R sumabs = R(0), prod = R(1), lg = R(0);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    SMath<R>::unit(xr[k], xi[k], sumabs, prod);
    if ((k & 3) == 3 || k == K - 1) lg += SMath<R>::flush(prod);
  }
  return SMath<R>::finish(sumabs, lg);
}

template <typename R, int L, int K, int C>
__global__ void __launch_bounds__(SAMP_WARPS * 32) sample_plain_kernel(NetDev<R> net, SampArgs a) {
  constexpr int G = 32 / L;      // chain groups per warp
  constexpr int NCH = G * C;     // chains per warp
  constexpr int BLK = 32 / NCH;  // Philox blocks per chain per RNG batch
  constexpr int SB = 4 * BLK;    // MH steps per RNG batch
  extern __shared__ uint32_t sig_all[];
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / L, l = lane % L;
  uint32_t *sig = sig_all + (size_t)warp * NCH * W32;  // [NCH][W32]
  const int64_t warp_global = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  const int64_t chain0 = warp_global * NCH;  // first local chain of this warp
  if (chain0 >= a.n_chains) return;
  const R s0 = (R)a.s0, s1 = (R)a.s1, ssum = s0 + s1;
  const RowLoader<R, L, K> rows(net.W, M, l);

  // ---- load / initialise configurations
  for (int idx = lane; idx < NCH * W32; idx += 32) {
    const int slot = idx / W32, w = idx - slot * W32;
    const int64_t ch = chain0 + slot;
    uint32_t v = 0;
    if (ch < a.n_chains) {
      if (a.init_random) {
        u32x4 ctr = {(uint32_t)(w >> 2), 0u, (uint32_t)(a.chain_offset + ch), (uint32_t)TAG_INIT};
        v = sel4(philox4x32_10(ctr, a.k0, a.k1), w & 3);
      } else {
        v = a.chain_state[ch * W32 + w];
      }
      const int nb = N - 32 * w;
      if (nb < 32) v &= (1u << nb) - 1u;
    }
    sig[idx] = v;
  }
  __syncwarp();

  // ---- theta from scratch: theta_j = b_j + sum_i x_i W_ij
  R thr[C][K], thi[C][K];
  {
    R br[K], bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int j = l + L * k;
      br[k] = (j < M && net.b) ? net.b[2 * j] : R(0);
      bi[k] = (j < M && net.b) ? net.b[2 * j + 1] : R(0);
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        thr[c][k] = br[k];
        thi[c][k] = bi[k];
      }
  }
  for (int i = 0; i < N; ++i) {
    R wr[K], wi[K];
    rows.load(i, wr, wi);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const R xv = ((sig[(g * C + c) * W32 + (i >> 5)] >> (i & 31)) & 1u) ? s1 : s0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        thr[c][k] = fma(xv, wr[k], thr[c][k]);
        thi[c][k] = fma(xv, wi[k], thi[c][k]);
      }
    }
  }
  R lc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) lc[c] = group_sum<L>(lane_logcosh2<R, K>(thr[c], thi[c]));

  unsigned n_acc = 0;
  const uint32_t n_steps = (uint32_t)(a.n_discard + a.chain_length) * (uint32_t)a.sweep_size;
  const uint32_t phase = (uint32_t)(a.step_offset & (uint64_t)(SB - 1));  // position of step 0 inside its batch
  const uint64_t batch0 = a.step_offset - phase;
  u32x4 rs = {0, 0, 0, 0};
  R lu[4] = {R(0), R(0), R(0), R(0)};
  int until_emit = a.sweep_size, sweep = 0;
  const bool shared_sites = (a.site_mode == 0);

  for (uint32_t s = 0; s < n_steps; ++s) {
    const uint32_t pos = s + phase;           // steps since batch0
    const int tb = (int)(pos & (SB - 1));     // position inside the RNG batch
    if (tb == 0 || s == 0) {
      // refill: lane r serves chain slot r % NCH with Philox block (batch start)/4 + r / NCH
      const uint64_t bb = batch0 + (uint64_t)(pos - tb);
      const uint64_t blk = (bb >> 2) + (uint64_t)(lane / NCH);
      const uint32_t chid = (uint32_t)(a.chain_offset + chain0 + (lane % NCH));
      u32x4 ctr = {(uint32_t)blk, (uint32_t)(blk >> 32), chid, (uint32_t)TAG_UNIFORM};
      const u32x4 ru = philox4x32_10(ctr, a.k0, a.k1);
      lu[0] = SMath<R>::log2u(ru.x);
      lu[1] = SMath<R>::log2u(ru.y);
      lu[2] = SMath<R>::log2u(ru.z);
      lu[3] = SMath<R>::log2u(ru.w);
      if (shared_sites) {
        const uint64_t sblk = (bb >> 2) + (uint64_t)(lane % BLK);
        u32x4 c2 = {(uint32_t)sblk, (uint32_t)(sblk >> 32), 0u, (uint32_t)TAG_SITE_SHARED};
        rs = philox4x32_10(c2, a.k0, a.k1);
      } else {
        ctr.w = (uint32_t)TAG_SITE_CHAIN;
        rs = philox4x32_10(ctr, a.k0, a.k1);
      }
    }
    const int bi = tb >> 2, comp = tb & 3;
    const uint32_t my_rs = sel4(rs, comp);
    const R my_lu = comp == 0 ? lu[0] : (comp == 1 ? lu[1] : (comp == 2 ? lu[2] : lu[3]));

    R wr[K], wi[K];
    int site = 0;
    R ar = R(0);
    if (shared_sites) {
      site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, bi), (uint32_t)N);
      rows.load(site, wr, wi);
      if (net.a) ar = __ldg(net.a + 2 * site);
    }

#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int slot = g * C + c;
      const int src = slot + NCH * bi;
      const R logu = __shfl_sync(0xffffffffu, my_lu, src);
      if (!shared_sites) {
        site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, src), (uint32_t)N);
        rows.load(site, wr, wi);
        ar = net.a ? __ldg(net.a + 2 * site) : R(0);
      }
      const uint32_t word = sig[slot * W32 + (site >> 5)];
      const R xold = ((word >> (site & 31)) & 1u) ? s1 : s0;
      const R d = ssum - R(2) * xold;  // x_new - x_old
      R pr[K], pi[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        pr[k] = fma(d, wr[k], thr[c][k]);
        pi[k] = fma(d, wi[k], thi[c][k]);
      }
      const R lcn = group_sum<L>(lane_logcosh2<R, K>(pr, pi));
      // log2 of the acceptance ratio |psi(x')|^2 / |psi(x)|^2
      const R dl = R(2) * (lcn - lc[c]) + R(2.0 * LOG2E) * d * ar;
      const bool acc = (logu < dl) && (chain0 + slot < a.n_chains);
      const R de = acc ? d : R(0);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        thr[c][k] = fma(de, wr[k], thr[c][k]);
        thi[c][k] = fma(de, wi[k], thi[c][k]);
      }
      lc[c] = acc ? lcn : lc[c];
      if (acc && l == 0) {
        sig[slot * W32 + (site >> 5)] = word ^ (1u << (site & 31));
        ++n_acc;
      }
    }
    __syncwarp();

    if (--until_emit == 0) {
      until_emit = a.sweep_size;
      if (sweep >= a.n_discard) {
        const int ts = sweep - a.n_discard;
        for (int idx = lane; idx < NCH * W32; idx += 32) {
          const int slot = idx / W32, w = idx - slot * W32;
          const int64_t ch = chain0 + slot;
          if (ch < a.n_chains) a.samples[(ch * a.chain_length + ts) * W32 + w] = sig[idx];
        }
      }
      ++sweep;
    }
  }
  for (int idx = lane; idx < NCH * W32; idx += 32) {
    const int slot = idx / W32, w = idx - slot * W32;
    const int64_t ch = chain0 + slot;
    if (ch < a.n_chains) a.chain_state[ch * W32 + w] = sig[idx];
  }
  if (a.n_accepted) {
    n_acc = warp_sum(n_acc);
    if (lane == 0 && n_acc) atomicAdd(a.n_accepted, (unsigned long long)n_acc);
  }
}

template <typename R, int L, int K, int C>
static int launch_plain(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, cudaStream_t st) {
  constexpr int NCH = (32 / L) * C;
  const int W32 = (net->n_visible + 31) >> 5;
  const int64_t warps = (a.n_chains + NCH - 1) / NCH;
  const int64_t blocks = (warps + SAMP_WARPS - 1) / SAMP_WARPS;
  const size_t smem = (size_t)SAMP_WARPS * NCH * W32 * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    cudaError_t rc = cudaFuncSetAttribute(sample_plain_kernel<R, L, K, C>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "sampler needs %zu B of shared memory", smem);
  }
  sample_plain_kernel<R, L, K, C><<<(unsigned)blocks, SAMP_WARPS * 32, smem, st>>>(make_netdev<R>(net), a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
