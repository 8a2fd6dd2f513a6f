// Metropolis-local sampler for a plain RBM target |psi|^2 (see sampler.cu for the overview).
//
// Template parameters: R real type; L lanes per chain group (4/8/16/32); K hidden units per
// lane, EXACT (K = ceil(M / L): no per-unit predicates in the Metropolis loop, only the last
// unit of a lane can fall outside M and it is handled by one predicated load);
// C chains per lane group (they share the row W[i, :] when the site sequence is shared).
//
// The kernel works on PRE-SCALED parameters written by prep_sampler_params into a scratch
// buffer:  X = 2 log2(e) Re(theta),  Y = 2 Im(theta), so that per hidden unit
//     log2 |cosh theta|^2 + 2 = |X| + log2(1 + t^2 + 2 t cos Y),   t = 2^(-|X|)
// costs MUFU.EX2 + MUFU.COS (+ MUFU.LG2 per four units: the log arguments are multiplied first)
// and, on sm_100, five packed-FP32 instructions per PAIR of units (FFMA2 / FMUL2: the (re, im)
// pair of a unit and the pair (unit k, unit k+1) are both natural float2 operands).
#pragma once

#include "sampler_common.cuh"

namespace nkfb {

// ------------------------------------------------------------------ packed pair arithmetic
template <typename R>
struct P2;
template <>
struct P2<float> {
  using V = float2;
  static __device__ __forceinline__ V mk(float a, float b) { return make_float2(a, b); }
  static __device__ __forceinline__ V fma(V a, V b, V c) { return __ffma2_rn(a, b, c); }
  static __device__ __forceinline__ V mul(V a, V b) { return __fmul2_rn(a, b); }
  static __device__ __forceinline__ float ex2(float x) { return mufu_ex2(x); }
  static __device__ __forceinline__ float cosr(float x) { return mufu_cos(x); }
  static __device__ __forceinline__ float lg2(float x) { return mufu_lg2(x); }
};
template <>
struct P2<double> {
  using V = double2;
  static __device__ __forceinline__ V mk(double a, double b) { return make_double2(a, b); }
  static __device__ __forceinline__ V fma(V a, V b, V c) { return make_double2(::fma(a.x, b.x, c.x), ::fma(a.y, b.y, c.y)); }
  static __device__ __forceinline__ V mul(V a, V b) { return make_double2(a.x * b.x, a.y * b.y); }
  static __device__ __forceinline__ double ex2(double x) { return exp2(x); }
  static __device__ __forceinline__ double cosr(double x) { return cos(x); }
  static __device__ __forceinline__ double lg2(double x) { return log2(x); }
};

__device__ __forceinline__ float abs_(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }

// scratch layout (reals of type R): Wt (N*M pairs) | bt (M pairs) | ar2 (N)
inline size_t sampler_ws_reals(int N, int M) { return 2 * (size_t)N * M + 2 * (size_t)M + (size_t)N; }

template <typename R>
__global__ void prep_sampler_params(NetDev<R> net, R *ws) {
  const int N = net.N, M = net.M;
  const int64_t nW = (int64_t)N * M;
  const R sx = (R)(2.0 * LOG2E);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nW + M + N;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < nW) {
      ws[2 * i] = sx * net.W[2 * i];
      ws[2 * i + 1] = R(2) * net.W[2 * i + 1];
    } else if (i < nW + M) {
      const int64_t j = i - nW;
      ws[2 * i] = net.b ? sx * net.b[2 * j] : R(0);
      ws[2 * i + 1] = net.b ? R(2) * net.b[2 * j + 1] : R(0);
    } else {
      const int64_t k = i - nW - M;
      ws[2 * (nW + M) + k] = net.a ? sx * net.a[2 * k] : R(0);
    }
  }
}

// sum over the lane's K units of |X| + log2(1 + t^2 + 2 t cos Y)
template <typename R, int K>
__device__ __forceinline__ R lane_lc(const typename P2<R>::V (&p)[K]) {
  using P = P2<R>;
  using V = typename P::V;
  R sumabs = R(0), lg = R(0);
  V prod = P::mk(R(1), R(1));
#pragma unroll
  for (int k = 0; k + 1 < K; k += 2) {
    const R a0 = abs_(p[k].x), a1 = abs_(p[k + 1].x);
    const V T = P::mk(P::ex2(-a0), P::ex2(-a1));
    const V Cc = P::mk(P::cosr(p[k].y), P::cosr(p[k + 1].y));
    const V u = P::fma(P::mk(R(2), R(2)), Cc, T);
    const V s = P::fma(T, u, P::mk(R(1), R(1)));
    prod = P::mul(prod, s);
    sumabs += a0;
    sumabs += a1;
    if (((k + 2) & 3) == 0) {  // every four units: one log2 for the product of four arguments
      lg += P::lg2(prod.x * prod.y);
      prod = P::mk(R(1), R(1));
    }
  }
  if (K & 1) {
    const R a0 = abs_(p[K - 1].x);
    const R t = P::ex2(-a0), c = P::cosr(p[K - 1].y);
    prod.x *= ::fma(t, ::fma(R(2), c, t), R(1));
    sumabs += a0;
  }
  if ((K & 3) != 0) lg += P::lg2(prod.x * prod.y);
  return sumabs + lg;
}

template <typename R, int L, int K>
struct RowLoader {
  using V = typename P2<R>::V;
  const V *W;  // pre-scaled complex rows, M per row
  int M, l;
  bool last_ok;
  __device__ __forceinline__ RowLoader(const R *Wp, int M_, int l_) : W(reinterpret_cast<const V *>(Wp)), M(M_), l(l_) {
    last_ok = (l + L * (K - 1)) < M;
  }
  __device__ __forceinline__ const V *ptr(int row) const { return W + (int64_t)row * M + l; }
  template <int k>
  __device__ __forceinline__ V ld(const V *p) const {
    if (k < K - 1) return __ldg(p + L * k);
    return last_ok ? __ldg(p + L * (K - 1)) : P2<R>::mk(R(0), R(0));
  }
  __device__ __forceinline__ void load(int row, V (&w)[K]) const {
    const V *p = W + (int64_t)row * M + l;
#pragma unroll
    for (int k = 0; k < K - 1; ++k) w[k] = __ldg(p + L * k);
    V v = P2<R>::mk(R(0), R(0));
    if (last_ok) v = __ldg(p + L * (K - 1));
    w[K - 1] = v;
  }
};

// Same sum as lane_lc for the PROPOSAL theta + d W[site, :], with the row streamed from L1/L2 instead of
// held in registers (used when K is large: the row is re-streamed for the accepted update).
template <typename R, int L, int K, int k>
struct StreamStep {
  using P = P2<R>;
  using V = typename P::V;
  static __device__ __forceinline__ void run(const RowLoader<R, L, K> &rows, const V *rp, V d2, const V (&th)[K],
                                             R &sumabs, R &lg, V &prod) {
    if constexpr (k + 1 < K) {
      const V p0 = P::fma(d2, rows.template ld<k>(rp), th[k]);
      const V p1 = P::fma(d2, rows.template ld<k + 1>(rp), th[k + 1]);
      const R a0 = abs_(p0.x), a1 = abs_(p1.x);
      const V T = P::mk(P::ex2(-a0), P::ex2(-a1));
      const V Cc = P::mk(P::cosr(p0.y), P::cosr(p1.y));
      const V u = P::fma(P::mk(R(2), R(2)), Cc, T);
      prod = P::mul(prod, P::fma(T, u, P::mk(R(1), R(1))));
      sumabs += a0;
      sumabs += a1;
      if (((k + 2) & 3) == 0) {
        lg += P::lg2(prod.x * prod.y);
        prod = P::mk(R(1), R(1));
      }
      StreamStep<R, L, K, k + 2>::run(rows, rp, d2, th, sumabs, lg, prod);
    } else if constexpr (k < K) {  // odd tail
      const V p0 = P::fma(d2, rows.template ld<k>(rp), th[k]);
      const R a0 = abs_(p0.x);
      const R t = P::ex2(-a0), c = P::cosr(p0.y);
      prod.x *= ::fma(t, ::fma(R(2), c, t), R(1));
      sumabs += a0;
    }
  }
};

template <typename R, int L, int K>
__device__ __forceinline__ R lane_lc_stream(const RowLoader<R, L, K> &rows, const typename P2<R>::V *rp,
                                            typename P2<R>::V d2, const typename P2<R>::V (&th)[K]) {
  using P = P2<R>;
  R sumabs = R(0), lg = R(0);
  typename P::V prod = P::mk(R(1), R(1));
  StreamStep<R, L, K, 0>::run(rows, rp, d2, th, sumabs, lg, prod);
  if ((K & 3) != 0) lg += P::lg2(prod.x * prod.y);
  return sumabs + lg;
}

template <typename R, int L, int K, int k>
struct StreamUpdate {
  using V = typename P2<R>::V;
  static __device__ __forceinline__ void run(const RowLoader<R, L, K> &rows, const V *rp, V de2, V (&th)[K]) {
    if constexpr (k < K) {
      th[k] = P2<R>::fma(de2, rows.template ld<k>(rp), th[k]);
      StreamUpdate<R, L, K, k + 1>::run(rows, rp, de2, th);
    }
  }
};

template <typename R, int L, int K, int C, int MB, bool STREAM>
__global__ void __launch_bounds__(SAMP_WARPS * 32, MB) sample_plain_kernel(NetDev<R> net, const R *ws, const R s0,
                                                                            const R s1, SampArgs a) {
  using P = P2<R>;
  using V = typename P::V;
  constexpr int G = 32 / L;      // chain groups per warp
  constexpr int NCH = G * C;     // chains per warp
  constexpr int BLK = 32 / NCH;  // Philox blocks per chain per RNG batch
  constexpr int SB = 4 * BLK;    // MH steps per RNG batch
  extern __shared__ uint32_t sig_all[];
  if (!samp_window_ok(a)) return;  // another kernel of the same call covers this parameter range
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / L, l = lane % L;
  uint32_t *sig = sig_all + (size_t)warp * NCH * W32;  // [NCH][W32]
  const int64_t warp_global = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  const int64_t chain0 = warp_global * NCH;  // first local chain of this warp
  if (chain0 >= a.n_chains) return;
  const R ssum = s0 + s1;
  const RowLoader<R, L, K> rows(ws, M, l);
  const V *bt = reinterpret_cast<const V *>(ws) + (size_t)N * M;
  const R *ar2 = ws + 2 * ((size_t)N * M + M);

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

  // ---- (scaled) theta from scratch: theta_j = b_j + sum_i x_i W_ij
  V th[C][K];
  {
    V bb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int j = l + L * k;
      bb[k] = (j < M) ? bt[j] : P::mk(R(0), R(0));
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int k = 0; k < K; ++k) th[c][k] = bb[k];
  }
  for (int i = 0; i < N; ++i) {
    V w[K];
    rows.load(i, w);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const R xv = ((sig[(g * C + c) * W32 + (i >> 5)] >> (i & 31)) & 1u) ? s1 : s0;
      const V xx = P::mk(xv, xv);
#pragma unroll
      for (int k = 0; k < K; ++k) th[c][k] = P::fma(xx, w[k], th[c][k]);
    }
  }
  R lc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) lc[c] = group_sum<L>(lane_lc<R, K>(th[c]));

  unsigned n_acc = 0;
  const uint32_t n_steps = (uint32_t)(a.n_discard + a.chain_length) * (uint32_t)a.sweep_size;
  const uint32_t phase = (uint32_t)(a.step_offset & (uint64_t)(SB - 1));  // position of step 0 inside its batch
  const uint64_t batch0 = a.step_offset - phase;
  u32x4 rs = {0, 0, 0, 0};
  R lu[4] = {R(0), R(0), R(0), R(0)};
  int until_emit = a.sweep_size, sweep = 0;
  const bool shared_sites = (a.site_mode == 0);
  const R invL = R(1) / R(L);

  for (uint32_t s = 0; s < n_steps; ++s) {
    const uint32_t pos = s + phase;        // steps since batch0
    const int tb = (int)(pos & (SB - 1));  // position inside the RNG batch
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

    if constexpr (STREAM) {
      // ---- phase 1: the row W[site, :] is streamed through registers (L1-resident), nothing is kept
      R part[C], dd[C], logu[C], arc[C];
      uint32_t word[C];
      int sitec[C];
      int site = 0;
      if (shared_sites) site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, bi), (uint32_t)N);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int slot = g * C + c;
        const int src = slot + NCH * bi;
        logu[c] = __shfl_sync(0xffffffffu, my_lu, src);
        if (!shared_sites) site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, src), (uint32_t)N);
        sitec[c] = site;
        arc[c] = __ldg(ar2 + site);
        word[c] = sig[slot * W32 + (site >> 5)];
        const R xold = ((word[c] >> (site & 31)) & 1u) ? s1 : s0;
        const R d = ssum - R(2) * xold;
        dd[c] = d;
        part[c] = ::fma(d * arc[c], invL, lane_lc_stream<R, L, K>(rows, rows.ptr(site), P::mk(d, d), th[c]));
      }
      // ---- phase 2: butterfly sums (every lane group of the warp reduces its own chain in the same shuffles)
#pragma unroll
      for (int m = L / 2; m > 0; m >>= 1)
#pragma unroll
        for (int c = 0; c < C; ++c) part[c] += __shfl_xor_sync(0xffffffffu, part[c], m);
      // ---- phase 3: accept / reject; the row is re-streamed only if some chain of the warp moved
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int slot = g * C + c;
        const R d = dd[c];
        const bool acc = (logu[c] < part[c] - lc[c]) && (chain0 + slot < a.n_chains);
        if (__any_sync(0xffffffffu, acc)) {
          const R de = acc ? d : R(0);
          StreamUpdate<R, L, K, 0>::run(rows, rows.ptr(sitec[c]), P::mk(de, de), th[c]);
        }
        lc[c] = acc ? part[c] - d * arc[c] : lc[c];
        if (acc && l == 0) {
          sig[slot * W32 + (sitec[c] >> 5)] = word[c] ^ (1u << (sitec[c] & 31));
          ++n_acc;
        }
      }
    } else {
      // ---- proposal site and the row W[site, :]
      V w[K];
      int site = 0;
      R ar = R(0);
      if (shared_sites) {
        site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, bi), (uint32_t)N);
        rows.load(site, w);
        ar = __ldg(ar2 + site);
      }

      // ---- phase 1: per-lane partial sums for every chain of the group
      R part[C], dd[C], logu[C];
      uint32_t word[C];
      int sitec[C];
  #pragma unroll
      for (int c = 0; c < C; ++c) {
        const int slot = g * C + c;
        const int src = slot + NCH * bi;
        logu[c] = __shfl_sync(0xffffffffu, my_lu, src);
        if (!shared_sites) {
          site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, src), (uint32_t)N);
          rows.load(site, w);
          ar = __ldg(ar2 + site);
        }
        sitec[c] = site;
        word[c] = sig[slot * W32 + (site >> 5)];
        const R xold = ((word[c] >> (site & 31)) & 1u) ? s1 : s0;
        const R d = ssum - R(2) * xold;  // x_new - x_old
        dd[c] = d;
        const V d2 = P::mk(d, d);
        V p[K];
  #pragma unroll
        for (int k = 0; k < K; ++k) p[k] = P::fma(d2, w[k], th[c][k]);
        // the visible-bias term d * ar is spread over the L lanes so that the butterfly sum carries it
        part[c] = ::fma(d * ar, invL, lane_lc<R, K>(p));
        if (!shared_sites) {
          // per-chain sites: the row differs from chain to chain, so finish this chain now
          const R lcn = group_sum<L>(part[c]);
          const bool acc = (logu[c] < lcn - lc[c]) && (chain0 + slot < a.n_chains);
          const R de = acc ? d : R(0);
          const V de2 = P::mk(de, de);
  #pragma unroll
          for (int k = 0; k < K; ++k) th[c][k] = P::fma(de2, w[k], th[c][k]);
          lc[c] = acc ? lcn - d * ar : lc[c];
          if (acc && l == 0) {
            sig[slot * W32 + (site >> 5)] = word[c] ^ (1u << (site & 31));
            ++n_acc;
          }
        }
      }
      if (shared_sites) {
        // ---- phase 2: the C butterfly reductions are independent and pipeline through the shuffle unit
  #pragma unroll
        for (int m = L / 2; m > 0; m >>= 1)
  #pragma unroll
          for (int c = 0; c < C; ++c) part[c] += __shfl_xor_sync(0xffffffffu, part[c], m);
        // ---- phase 3: accept iff log2 u < log2 |psi(x')|^2 / |psi(x)|^2
  #pragma unroll
        for (int c = 0; c < C; ++c) {
          const int slot = g * C + c;
          const R d = dd[c];
          const bool acc = (logu[c] < part[c] - lc[c]) && (chain0 + slot < a.n_chains);
          const R de = acc ? d : R(0);
          const V de2 = P::mk(de, de);
  #pragma unroll
          for (int k = 0; k < K; ++k) th[c][k] = P::fma(de2, w[k], th[c][k]);
          lc[c] = acc ? part[c] - d * ar : lc[c];
          if (acc && l == 0) {
            sig[slot * W32 + (sitec[c] >> 5)] = word[c] ^ (1u << (sitec[c] & 31));
            ++n_acc;
          }
        }
      }
    }
    __syncwarp();

    if (--until_emit == 0) {
      until_emit = a.sweep_size;
      if (sweep >= a.n_discard) {
        const int ts = sweep - a.n_discard;
        for (int idx = lane; idx < NCH * W32; idx += 32) {
          const int slot = idx / W32, w32 = idx - slot * W32;
          const int64_t ch = chain0 + slot;
          if (ch < a.n_chains) a.samples[(ch * a.chain_length + ts) * W32 + w32] = sig[idx];
        }
      }
      ++sweep;
    }
  }
  for (int idx = lane; idx < NCH * W32; idx += 32) {
    const int slot = idx / W32, w32 = idx - slot * W32;
    const int64_t ch = chain0 + slot;
    if (ch < a.n_chains) a.chain_state[ch * W32 + w32] = sig[idx];
  }
  if (a.n_accepted) {
    n_acc = warp_sum(n_acc);
    if (lane == 0 && n_acc) atomicAdd(a.n_accepted, (unsigned long long)n_acc);
  }
}

template <typename R, int L, int K, int C, int MB = 1, bool STREAM = false>
static int launch_plain(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, const void *ws, cudaStream_t st) {
  constexpr int NCH = (32 / L) * C;
  const int W32 = (net->n_visible + 31) >> 5;
  const int64_t warps = (a.n_chains + NCH - 1) / NCH;
  const int64_t blocks = (warps + SAMP_WARPS - 1) / SAMP_WARPS;
  const size_t smem = (size_t)SAMP_WARPS * NCH * W32 * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    cudaError_t rc = cudaFuncSetAttribute(sample_plain_kernel<R, L, K, C, MB, STREAM>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "sampler needs %zu B of shared memory", smem);
  }
  sample_plain_kernel<R, L, K, C, MB, STREAM><<<(unsigned)blocks, SAMP_WARPS * 32, smem, st>>>(
      make_netdev<R>(net), (const R *)ws, (R)a.s0, (R)a.s1, a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
