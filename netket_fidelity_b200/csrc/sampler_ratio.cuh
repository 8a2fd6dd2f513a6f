// Metropolis-local sampler for a plain RBM target |psi|^2, MULTIPLICATIVE form ("ratio sampler").
//
// Instead of the hidden pre-activations theta_j the chain keeps  z_j = exp(-2 theta_j)  (complex).
// A single-spin flip at site i with x_new - x_old = d multiplies it by a constant of the parameters,
//     z'_j = z_j * q_ij,   q_ij = exp(-2 d W_ij)          (two tables: d = +(s1 - s0) and d = -(s1 - s0)),
// and because  cosh(theta) = e^theta (1 + z) / 2  the acceptance ratio needs no transcendental per unit:
//     log2 |psi(x')|^2 - log2 |psi(x)|^2 = A_i(d) + sum_j [ log2 |1 + z'_j|^2 - log2 |1 + z_j|^2 ],
//     A_i(d) = 2 log2(e) d (Re a_i + sum_j Re W_ij).
// Per hidden unit and proposal: one complex multiply, |1 + z'|^2 and one factor of a running product --
// on sm_100 eight packed-FP32 instructions (FMUL2 / FFMA2 / FADD2) per PAIR of units -- and one MUFU.LG2 per
// GL pairs (the logs of 2 GL units are taken from their product).  The kernel is bound by the FP32 pipe and
// by the L1 data path that streams the q rows (16 B per pair and chain), not by the MUFU pipe as the additive
// sampler (sampler_plain.cuh: EX2 + COS per unit) is.
//
// Range: a product of 2 GL factors |1 + z|^2 <= (1 + e^{2X})^2 must stay below 2^126, X = max_j (|Re b_j| +
// max|s| sum_i |Re W_ij|) >= |Re theta_j| for every configuration.  prep_ratio_params writes X next to the
// tables; every sampler kernel starts by comparing X with its own window (SampArgs::xlo, xhi) and returns
// immediately when it is not the one that covers X, so the host launches [ratio GL=4 | ratio GL=1 | additive]
// back to back without reading anything back and exactly one of them does the work.
//
// Random numbers, site sequences, layouts and the emitted samples follow sampler_plain.cuh exactly (the two
// kernels accept the same proposals up to rounding of the log-ratio).
//
// Template parameters: R real type; L lanes per chain group (4/8/16/32); NP PAIRS of hidden units per lane
// (pair p of lane l = units 2 L p + l and 2 L p + L + l); C chains per lane group; GL pairs per logarithm;
// MB minimum resident CTAs per SM.
#pragma once

#include "sampler_plain.cuh"

namespace nkfb {

template <typename R>
struct V4T;
template <>
struct V4T<float> {
  float2 lo, hi;
  static __device__ __forceinline__ V4T<float> ld(const float *p) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    V4T<float> r;
    r.lo = make_float2(v.x, v.y);
    r.hi = make_float2(v.z, v.w);
    return r;
  }
};
template <>
struct V4T<double> {
  double2 lo, hi;
  static __device__ __forceinline__ V4T<double> ld(const double *p) {
    V4T<double> r;
    r.lo = __ldg(reinterpret_cast<const double2 *>(p));
    r.hi = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return r;
  }
};

__device__ __forceinline__ float2 add2_(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ double2 add2_(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 neg2_(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ double2 neg2_(double2 a) { return make_double2(-a.x, -a.y); }

// ---- workspace layout (reals of type R):  Q [2][N][NP * L][4]  |  A0 [N]  |  pad to 4  |  X (float, 1 word)
inline int ratio_pairs(int M, int L) { return (M + 2 * L - 1) / (2 * L); }
inline size_t ratio_q_reals(int N, int M, int L) { return (size_t)8 * N * ratio_pairs(M, L) * L; }
inline size_t ratio_a0_off(int N, int M, int L) { return ratio_q_reals(N, M, L); }
inline size_t ratio_bound_off(int N, int M, int L) { return (ratio_q_reals(N, M, L) + (size_t)N + 3) & ~(size_t)3; }
// upper bound over every lane-group width (NP * L <= M / 2 + 32)
inline size_t ratio_ws_reals_max(int N, int M) { return (size_t)8 * N * ((M + 1) / 2 + 32) + (size_t)N + 8; }

// largest X for which a product of 2 GL factors (1 + e^{2X})^2 stays below 2^126 (2 % margin)
inline float ratio_xmax(int GL) {
  const double lim = 126.0 * LN2 / (4.0 * GL);  // ln(1 + e^{2X}) < lim
  return (float)(0.98 * 0.5 * log(expm1(lim)));
}

template <typename R>
__global__ void prep_ratio_params(NetDev<R> net, R *ws, int L, int NP, R s0, R s1, size_t a0_off, size_t bound_off,
                                  int tables_only = 0) {
  const int N = net.N, M = net.M;
  const int64_t slots_per_tab = (int64_t)N * NP * L;
  const int64_t n_slots = 2 * slots_per_tab;
  const double dx = (double)s1 - (double)s0;
  // the N + M reductions (A_i over the hidden units, the range bound over the sites) take one WARP each: one thread each
  // made this kernel 29 us long (400 dependent loads), on the critical path of every sampler launch
  // (the reduction items start at a warp boundary: a warp is either all table slots or all lanes of one item)
  const int64_t red0 = (n_slots + 31) & ~(int64_t)31;
  const int64_t n_red = tables_only ? 0 : (int64_t)(N + M) * 32;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < red0 + n_red;
       idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx < n_slots) {
      const int tab = idx >= slots_per_tab;
      const int64_t rem = idx - tab * slots_per_tab;
      const int i = (int)(rem / (NP * L));
      const int pl = (int)(rem - (int64_t)i * NP * L);
      const int p = pl / L, l = pl - p * L;
      const R sgn = (R)(tab == 0 ? -2.0 * dx : 2.0 * dx);  // table 0: flip s0 -> s1 (d = +dx), q = exp(-2 d W)
      R out[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = 2 * L * p + u * L + l;
        R qr = R(1), qi = R(0);
        if (j < M) {
          const R wr = net.W[2 * ((int64_t)i * M + j)], wi = net.W[2 * ((int64_t)i * M + j) + 1];
          const R e = exp_(sgn * wr);
          R sn, cs;
          sincos_(sgn * wi, &sn, &cs);
          qr = e * cs;
          qi = e * sn;
        }
        out[u] = qr;
        out[2 + u] = qi;
      }
      st4(ws + 4 * idx, out);
    } else if (idx >= red0) {
      const int64_t item = (idx - red0) >> 5;
      const int ln = (int)((idx - red0) & 31);
      if (item < N) {
        const int i = (int)item;
        double acc = 0.0;
        for (int j = ln; j < M; j += 32) acc += (double)net.W[2 * ((int64_t)i * M + j)];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (ln == 0) ws[a0_off + i] = (R)(2.0 * LOG2E * dx * (acc + (net.a ? (double)net.a[2 * i] : 0.0)));
      } else {
        const int j = (int)(item - N);
        const double smax = fmax(fabs((double)s0), fabs((double)s1));
        double acc = 0.0;
        for (int i = ln; i < N; i += 32) acc += fabs((double)net.W[2 * ((int64_t)i * M + j)]);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (ln == 0) {
          acc = acc * smax + (net.b ? fabs((double)net.b[2 * j]) : 0.0);
          float xb = __double2float_ru(acc);
          if (!(xb == xb)) xb = __int_as_float(0x7f800000);  // NaN parameters -> no fast path
          atomicMax(reinterpret_cast<unsigned int *>(ws + bound_off), __float_as_uint(xb));  // xb >= 0: bit order = value order
        }
      }
    }
  }
}

// sum over a lane group of L lanes.  Full warps use one integer REDUX on a 2^-20 fixed-point image of the
// addends (|v| <= 60 on every lane, else the butterfly): the reduction sits on the critical path of every
// Metropolis step and a butterfly costs five dependent shuffle + add rounds.
template <int L>
__device__ __forceinline__ float group_sum_fast(float v) {
  if constexpr (L == 32) {
    if (__all_sync(0xffffffffu, fabsf(v) <= 60.0f)) {
      const int s = __reduce_add_sync(0xffffffffu, __float2int_rn(v * 1048576.0f));
      return (float)s * (1.0f / 1048576.0f);
    }
  }
  return group_sum<L>(v);
}
template <int L>
__device__ __forceinline__ double group_sum_fast(double v) { return group_sum<L>(v); }

// MODE 0: the proposal z' is kept next to z and copied over it when the move is accepted.
// MODE 1: z is multiplied in place and multiplied back by the inverse row (the other table, same site) when the
//         move is REJECTED: half the registers, no copies; z is recomputed from the configuration every
//         RATIO_REFRESH sweeps, which bounds the rounding drift of the products (both modes).
constexpr int RATIO_REFRESH = 32;
constexpr int PIPE_WR_ = 4;  // configuration words kept in registers by the one-warp-per-chain kernels: N <= 128

template <typename R, int L, int NP, int C, int GL, int MB, int MODE>
__global__ void __launch_bounds__(SAMP_WARPS * 32, MB) sample_ratio_kernel(NetDev<R> net, const R *ws, const R s0,
                                                                            const R s1, SampArgs a) {
  using P = P2<R>;
  using V = typename P::V;
  using V4 = V4T<R>;
  constexpr int G = 32 / L;      // chain groups per warp
  constexpr int NCH = G * C;     // chains per warp
  constexpr int BLK = 32 / NCH;  // Philox blocks per chain per RNG batch
  constexpr int SB = 4 * BLK;    // MH steps per RNG batch
  extern __shared__ uint32_t sig_all[];
  if (!samp_window_ok(a)) return;
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / L, l = lane % L;
  uint32_t *sig = sig_all + (size_t)warp * NCH * W32;  // [NCH][W32]
  const int64_t warp_global = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  const int64_t chain0 = warp_global * NCH;  // first local chain of this warp
  if (chain0 >= a.n_chains) return;
  const int row_reals = 4 * NP * L;  // reals per (table, site) row
  const R *Q = ws + 4 * l;
  const R *A0 = ws + (size_t)8 * N * NP * L;
  const V one = P::mk(R(1), R(1));

  // ---- load / initialise configurations (identical to sample_plain_kernel)
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

  V Zr[C][NP], Zi[C][NP];
  R lg[C];  // per-LANE cache of sum_j log2 |1 + z_j|^2 over the lane's units (only differences are reduced)
  unsigned n_acc = 0;
  const uint32_t n_steps = (uint32_t)(a.n_discard + a.chain_length) * (uint32_t)a.sweep_size;
  const uint32_t phase = (uint32_t)(a.step_offset & (uint64_t)(SB - 1));  // position of step 0 inside its batch
  const uint64_t batch0 = a.step_offset - phase;
  u32x4 rs = {0, 0, 0, 0};
  R lu[4] = {R(0), R(0), R(0), R(0)};
  int until_emit = a.sweep_size, sweep = 0;
  const bool shared_sites = (a.site_mode == 0);
  const uint32_t seg_steps = (uint32_t)RATIO_REFRESH * (uint32_t)a.sweep_size;

  uint32_t s = 0;
  while (s < n_steps) {
    // ---- theta from scratch (pairs of units), then z = exp(-2 theta); padded units carry z = 0
    {
      const V *cW = reinterpret_cast<const V *>(net.W);
      const V *cb = reinterpret_cast<const V *>(net.b);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int j0 = 2 * L * p + l, j1 = j0 + L;
        const V b0 = (cb && j0 < M) ? cb[j0] : P::mk(R(0), R(0));
        const V b1 = (cb && j1 < M) ? cb[j1] : P::mk(R(0), R(0));
#pragma unroll
        for (int c = 0; c < C; ++c) {
          Zr[c][p] = P::mk(b0.x, b1.x);
          Zi[c][p] = P::mk(b0.y, b1.y);
        }
      }
      for (int i = 0; i < N; ++i) {
        V wr[NP], wi[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int j0 = 2 * L * p + l, j1 = j0 + L;
          const V w0 = (j0 < M) ? __ldg(cW + (int64_t)i * M + j0) : P::mk(R(0), R(0));
          const V w1 = (j1 < M) ? __ldg(cW + (int64_t)i * M + j1) : P::mk(R(0), R(0));
          wr[p] = P::mk(w0.x, w1.x);
          wi[p] = P::mk(w0.y, w1.y);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const R xv = ((sig[(g * C + c) * W32 + (i >> 5)] >> (i & 31)) & 1u) ? s1 : s0;
          const V xx = P::mk(xv, xv);
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            Zr[c][p] = P::fma(xx, wr[p], Zr[c][p]);
            Zi[c][p] = P::fma(xx, wi[p], Zi[c][p]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int j0 = 2 * L * p + l, j1 = j0 + L;
          R e0 = exp_(R(-2) * Zr[c][p].x), e1 = exp_(R(-2) * Zr[c][p].y);
          R sn0, cs0, sn1, cs1;
          sincos_(R(-2) * Zi[c][p].x, &sn0, &cs0);
          sincos_(R(-2) * Zi[c][p].y, &sn1, &cs1);
          if (j0 >= M) e0 = R(0);
          if (j1 >= M) e1 = R(0);
          Zr[c][p] = P::mk(e0 * cs0, e1 * cs1);
          Zi[c][p] = P::mk(e0 * sn0, e1 * sn1);
        }
        R acc = R(0);
        V prod = one;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const V aa = add2_(Zr[c][p], one);
          V m = P::mul(aa, aa);
          m = P::fma(Zi[c][p], Zi[c][p], m);
          prod = P::mul(prod, m);
          if (((p + 1) % GL) == 0 || p == NP - 1) {
            acc += P::lg2(prod.x * prod.y);
            prod = one;
          }
        }
        lg[c] = acc;
      }
    }

    const uint32_t s_end = (n_steps - s > seg_steps) ? s + seg_steps : n_steps;
    for (; s < s_end; ++s) {
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

      int site = 0;
      if (shared_sites) site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, bi), (uint32_t)N);

      // ---- phase 1: proposals z' = z q and the per-lane change of sum_j log2 |1 + z_j|^2
      V nZr[MODE == 0 ? C : 1][MODE == 0 ? NP : 1], nZi[MODE == 0 ? C : 1][MODE == 0 ? NP : 1];
      R part[C], lgn[C], logu[C], dA[C];
      uint32_t word[C];
      int sitec[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int slot = g * C + c;
        const int src = slot + NCH * bi;
        logu[c] = __shfl_sync(0xffffffffu, my_lu, src);
        if (!shared_sites) site = (int)__umulhi(__shfl_sync(0xffffffffu, my_rs, src), (uint32_t)N);
        sitec[c] = site;
        word[c] = sig[slot * W32 + (site >> 5)];
        const uint32_t bit = (word[c] >> (site & 31)) & 1u;
        const R *rp = Q + (size_t)((int)bit * N + site) * row_reals;
        const R a0 = __ldg(A0 + site);
        dA[c] = bit ? -a0 : a0;
        R acc = R(0);
        V prod = one;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const V4 q = V4::ld(rp + 4 * L * p);  // lo = (Re q_j0, Re q_j1), hi = (Im q_j0, Im q_j1)
          const V t = P::mul(Zi[c][p], q.hi);
          const V nr = P::fma(Zr[c][p], q.lo, neg2_(t));
          const V u = P::mul(Zi[c][p], q.lo);
          const V ni = P::fma(Zr[c][p], q.hi, u);
          if constexpr (MODE == 0) {
            nZr[c][p] = nr;
            nZi[c][p] = ni;
          } else {
            Zr[c][p] = nr;
            Zi[c][p] = ni;
          }
          const V aa = add2_(nr, one);
          V m = P::mul(aa, aa);
          m = P::fma(ni, ni, m);
          prod = P::mul(prod, m);
          if (((p + 1) % GL) == 0 || p == NP - 1) {
            acc += P::lg2(prod.x * prod.y);
            prod = one;
          }
        }
        lgn[c] = acc;
        part[c] = acc - lg[c];
      }
      // ---- phase 2: sums over the lane group
#pragma unroll
      for (int c = 0; c < C; ++c) part[c] = group_sum_fast<L>(part[c]);
      // ---- phase 3: accept iff log2 u < log2 |psi(x')|^2 - log2 |psi(x)|^2
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int slot = g * C + c;
        const bool acc = (logu[c] < part[c] + dA[c]) && (chain0 + slot < a.n_chains);
        if (acc) {
          if constexpr (MODE == 0) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              Zr[c][p] = nZr[c][p];
              Zi[c][p] = nZi[c][p];
            }
          }
          lg[c] = lgn[c];
          if (l == 0) {
            sig[slot * W32 + (sitec[c] >> 5)] = word[c] ^ (1u << (sitec[c] & 31));
            ++n_acc;
          }
        } else if constexpr (MODE == 1) {
          // rejected: multiply back by the inverse row (the row of the opposite flip at the same site)
          const uint32_t nbit = ((word[c] >> (sitec[c] & 31)) & 1u) ^ 1u;
          const R *rp = Q + (size_t)((int)nbit * N + sitec[c]) * row_reals;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const V4 q = V4::ld(rp + 4 * L * p);
            const V t = P::mul(Zi[c][p], q.hi);
            const V nr = P::fma(Zr[c][p], q.lo, neg2_(t));
            const V u = P::mul(Zi[c][p], q.lo);
            Zi[c][p] = P::fma(Zr[c][p], q.hi, u);
            Zr[c][p] = nr;
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

// ------------------------------------------------------------------------------------------------------------
// One chain per warp (L = 32), software-pipelined, copy-free.  The hot form of the multiplicative sampler.
//   * two register images of z: the multiply phase reads one and writes the other, an accepted move just
//     switches which one is current (no copy on accept, no multiply-back on reject);
//   * the q row of step s + 1 is requested right after the multiply phase of step s (the site sequence comes from
//     the RNG alone; the spin at that site is known unless step s flips the very same site, in which case the
//     prefetched row is inverted in place), so its L1 / L2 latency overlaps the logarithm - reduction -
//     decision tail of step s;
//   * a batch of 128 steps of random numbers (log2 u, proposal site, A_i) is staged in shared memory by one
//     Philox call per lane, each step reads its three words with a broadcast LDS;
//   * the multiply phase is written stage by stage over the NP independent pairs, products combined as trees.
// Random numbers -> steps exactly as in sample_plain_kernel / sample_ratio_kernel.
// Acceptance test of one chain: accept iff sum over the warp of `part` exceeds thr = log2 u - A.
// Single precision: the addends go through one integer REDUX as 2^-20 fixed point (a butterfly is five dependent
// shuffle + add rounds on the critical path of every step) and the threshold is compared as an integer; an addend
// outside +-60 (a hidden unit sitting exactly on a node of cosh: |1 + z|^2 underflows) saturates there, where the
// outcome no longer depends on it (log2 u >= -25).  Double precision: butterfly sum, floating compare.
template <typename R>
struct WarpDecide;
template <>
struct WarpDecide<float> {
  using T = int;
  static __device__ __forceinline__ T threshold(float thr) { return __float2int_rd(thr * 1048576.0f); }
  static __device__ __forceinline__ bool accept(float part, T thr) {
    const int v = __float2int_rn(fminf(fmaxf(part, -60.0f), 60.0f) * 1048576.0f);
    return __reduce_add_sync(0xffffffffu, v) > thr;
  }
};
template <>
struct WarpDecide<double> {
  using T = double;
  static __device__ __forceinline__ T threshold(double thr) { return thr; }
  static __device__ __forceinline__ bool accept(double part, T thr) { return group_sum<32>(part) > thr; }
};

// WR: configuration words kept in registers (N <= 32 WR).
template <typename R, int NP, int GL, int MB, int WR>
__global__ void __launch_bounds__(SAMP_WARPS * 32, MB) sample_ratio_pipe_kernel(NetDev<R> net, const R *ws,
                                                                                 const R s0, const R s1, SampArgs a) {
  using P = P2<R>;
  using V = typename P::V;
  using V4 = V4T<R>;
  using D = WarpDecide<R>;
  constexpr int L = 32;
  constexpr int SB = 128;                 // MH steps per RNG batch: 32 lanes x one Philox block of 4
  constexpr int NG = (NP + GL - 1) / GL;  // logarithms per step
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  if (!samp_window_ok(a)) return;
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = lane;
  constexpr size_t per_warp = (size_t)WR * 4 + (size_t)SB * (2 * sizeof(R) + 4);
  unsigned char *base = pipe_smem + (size_t)warp * per_warp;
  R *s_lu = reinterpret_cast<R *>(base);                   // [SB] log2 u
  R *s_a0 = s_lu + SB;                                     // [SB] A_site
  int *s_site = reinterpret_cast<int *>(s_a0 + SB);        // [SB] proposal site
  uint32_t *sig = reinterpret_cast<uint32_t *>(s_site + SB);  // [WR] configuration, synchronised once per sweep
  const int64_t chain = (int64_t)blockIdx.x * SAMP_WARPS + warp;  // local chain of this warp
  if (chain >= a.n_chains) return;
  const int row_reals = 4 * NP * L;  // reals per (table, site) row
  const R *Q = ws + 4 * l;
  const R *A0 = ws + (size_t)8 * N * NP * L;
  const V one = P::mk(R(1), R(1));

  // ---- configuration: WR words in shared memory (read with a broadcast LDS, flipped by lane 0; kept in registers
  // they turned into a local-memory array behind a jump table)
  for (int w = lane; w < WR; w += 32) {
    uint32_t v = 0;
    if (w < W32) {
      if (a.init_random) {
        u32x4 ctr = {(uint32_t)(w >> 2), 0u, (uint32_t)(a.chain_offset + chain), (uint32_t)TAG_INIT};
        v = sel4(philox4x32_10(ctr, a.k0, a.k1), w & 3);
      } else {
        v = a.chain_state[chain * W32 + w];
      }
      const int nb = N - 32 * w;
      if (nb < 32) v &= (1u << nb) - 1u;
    }
    sig[w] = v;
  }
  __syncwarp();
  auto spin_bit = [&](int st) -> uint32_t { return (sig[st >> 5] >> (st & 31)) & 1u; };
  auto sync_sig = [&]() { __syncwarp(); };

  const int n_sweeps = a.n_discard + a.chain_length;
  const uint32_t phase = (uint32_t)(a.step_offset & (uint64_t)(SB - 1));
  const uint64_t batch0 = a.step_offset - phase;
  const bool shared_sites = (a.site_mode == 0);
  const uint32_t chid = (uint32_t)(a.chain_offset + chain);

  // stage the random numbers of the batch that starts `pos0` steps after batch0 (pos0 a multiple of SB)
  auto refill = [&](uint32_t pos0) {
    const uint64_t blk = ((batch0 + (uint64_t)pos0) >> 2) + (uint64_t)lane;
    u32x4 ctr = {(uint32_t)blk, (uint32_t)(blk >> 32), chid, (uint32_t)TAG_UNIFORM};
    const u32x4 ru = philox4x32_10(ctr, a.k0, a.k1);
    u32x4 rsite;
    if (shared_sites) {
      u32x4 c2 = {(uint32_t)blk, (uint32_t)(blk >> 32), 0u, (uint32_t)TAG_SITE_SHARED};
      rsite = philox4x32_10(c2, a.k0, a.k1);
    } else {
      ctr.w = (uint32_t)TAG_SITE_CHAIN;
      rsite = philox4x32_10(ctr, a.k0, a.k1);
    }
    __syncwarp();
    const int i0 = (int)__umulhi(rsite.x, (uint32_t)N), i1 = (int)__umulhi(rsite.y, (uint32_t)N);
    const int i2 = (int)__umulhi(rsite.z, (uint32_t)N), i3 = (int)__umulhi(rsite.w, (uint32_t)N);
    s_lu[4 * lane + 0] = SMath<R>::log2u(ru.x);
    s_lu[4 * lane + 1] = SMath<R>::log2u(ru.y);
    s_lu[4 * lane + 2] = SMath<R>::log2u(ru.z);
    s_lu[4 * lane + 3] = SMath<R>::log2u(ru.w);
    s_site[4 * lane + 0] = i0;
    s_site[4 * lane + 1] = i1;
    s_site[4 * lane + 2] = i2;
    s_site[4 * lane + 3] = i3;
    s_a0[4 * lane + 0] = __ldg(A0 + i0);
    s_a0[4 * lane + 1] = __ldg(A0 + i1);
    s_a0[4 * lane + 2] = __ldg(A0 + i2);
    s_a0[4 * lane + 3] = __ldg(A0 + i3);
    __syncwarp();
  };

  V ZAr[NP], ZAi[NP], ZBr[NP], ZBi[NP];
  V4 row[NP];
  R lg = R(0);
  unsigned n_acc = 0, cur = 0;

  // ---- prologue: random numbers and row of step 0
  refill(0);
  int site = s_site[phase];
  uint32_t bit = spin_bit(site);
  R thr_r;
  {
    const R a0 = s_a0[phase];
    thr_r = s_lu[phase] - (bit ? -a0 : a0);
    const R *rp = Q + (size_t)((int)bit * N + site) * row_reals;
#pragma unroll
    for (int p = 0; p < NP; ++p) row[p] = V4::ld(rp + 4 * L * p);
  }
  typename D::T thr = D::threshold(thr_r);

  // multiply phase: zo = zi * row, returns this lane's sum_j log2 |1 + zo_j|^2
  auto multiply = [&](const V (&zir)[NP], const V (&zii)[NP], V (&zor)[NP], V (&zoi)[NP]) -> R {
    V t[NP], u[NP], m[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      t[p] = P::mul(zii[p], row[p].hi);
      u[p] = P::mul(zii[p], row[p].lo);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      zor[p] = P::fma(zir[p], row[p].lo, neg2_(t[p]));
      zoi[p] = P::fma(zir[p], row[p].hi, u[p]);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) m[p] = add2_(zor[p], one);
#pragma unroll
    for (int p = 0; p < NP; ++p) m[p] = P::mul(m[p], m[p]);
#pragma unroll
    for (int p = 0; p < NP; ++p) m[p] = P::fma(zoi[p], zoi[p], m[p]);
#pragma unroll
    for (int st = 1; st < GL; st <<= 1)
#pragma unroll
      for (int p = 0; p < NP; ++p)
        if ((p % GL) % (2 * st) == 0 && (p % GL) + st < GL && p + st < NP) m[p] = P::mul(m[p], m[p + st]);
    R acc = R(0);
#pragma unroll
    for (int gq = 0; gq < NG; ++gq) {
      const R lv = P::lg2(m[gq * GL].x * m[gq * GL].y);
      acc = gq == 0 ? lv : acc + lv;
    }
    return acc;
  };

  uint32_t s = 0;  // steps done
  for (int sweep = 0; sweep < n_sweeps; ++sweep) {
    if ((sweep % RATIO_REFRESH) == 0) {
      // ---- theta from scratch (pairs of units), then z = exp(-2 theta) into image A; padded units carry z = 0
      sync_sig();
      cur = 0;
      const V *cW = reinterpret_cast<const V *>(net.W);
      const V *cb = reinterpret_cast<const V *>(net.b);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int j0 = 2 * L * p + l, j1 = j0 + L;
        const V b0 = (cb && j0 < M) ? cb[j0] : P::mk(R(0), R(0));
        const V b1 = (cb && j1 < M) ? cb[j1] : P::mk(R(0), R(0));
        ZAr[p] = P::mk(b0.x, b1.x);
        ZAi[p] = P::mk(b0.y, b1.y);
      }
      for (int i = 0; i < N; ++i) {
        const R xv = ((sig[i >> 5] >> (i & 31)) & 1u) ? s1 : s0;
        const V xx = P::mk(xv, xv);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int j0 = 2 * L * p + l, j1 = j0 + L;
          const V w0 = (j0 < M) ? __ldg(cW + (int64_t)i * M + j0) : P::mk(R(0), R(0));
          const V w1 = (j1 < M) ? __ldg(cW + (int64_t)i * M + j1) : P::mk(R(0), R(0));
          ZAr[p] = P::fma(xx, P::mk(w0.x, w1.x), ZAr[p]);
          ZAi[p] = P::fma(xx, P::mk(w0.y, w1.y), ZAi[p]);
        }
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int j0 = 2 * L * p + l, j1 = j0 + L;
        R e0 = exp_(R(-2) * ZAr[p].x), e1 = exp_(R(-2) * ZAr[p].y);
        R sn0, cs0, sn1, cs1;
        sincos_(R(-2) * ZAi[p].x, &sn0, &cs0);
        sincos_(R(-2) * ZAi[p].y, &sn1, &cs1);
        if (j0 >= M) e0 = R(0);
        if (j1 >= M) e1 = R(0);
        ZAr[p] = P::mk(e0 * cs0, e1 * cs1);
        ZAi[p] = P::mk(e0 * sn0, e1 * sn1);
      }
      R acc = R(0);
#pragma unroll
      for (int gq = 0; gq < NG; ++gq) {
        V prod = one;
#pragma unroll
        for (int p = gq * GL; p < (gq + 1) * GL && p < NP; ++p) {
          const V aa = add2_(ZAr[p], one);
          V m = P::mul(aa, aa);
          m = P::fma(ZAi[p], ZAi[p], m);
          prod = P::mul(prod, m);
        }
        acc += P::lg2(prod.x * prod.y);
      }
      lg = acc;
    }

    for (int k = 0; k < a.sweep_size; ++k, ++s) {
      // ---- random numbers of step s + 1 (harmless one step past the end); address of its row
      const uint32_t npos = s + 1 + phase;
      const int ntb = (int)(npos & (SB - 1));
      if (ntb == 0) refill(npos);
      const int nsite = s_site[ntb];
      const R nlogu = s_lu[ntb], na0 = s_a0[ntb];
      uint32_t nbit = spin_bit(nsite);
      const R *nrp = Q + (size_t)((int)nbit * N + nsite) * row_reals;
      R nthr_r = nlogu - (nbit ? -na0 : na0);

      // ---- proposal z' = z q into the other image
      R lgn;
      if (cur == 0)
        lgn = multiply(ZAr, ZAi, ZBr, ZBi);
      else
        lgn = multiply(ZBr, ZBi, ZAr, ZAi);

      // ---- request the row of step s + 1: its latency overlaps the reduction and the decision
#pragma unroll
      for (int p = 0; p < NP; ++p) row[p] = V4::ld(nrp + 4 * L * p);

      // ---- accept iff log2 u < log2 |psi(x')|^2 - log2 |psi(x)|^2
      if (D::accept(lgn - lg, thr)) {
        lg = lgn;
        cur ^= 1u;
        ++n_acc;
        if (lane == 0) sig[site >> 5] ^= 1u << (site & 31);
        if (nsite == site) {
          // the next proposal flips the same spin back: its row is the inverse of the one requested above
          nthr_r = nlogu - (nbit ? na0 : -na0);
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const V d = P::fma(row[p].hi, row[p].hi, P::mul(row[p].lo, row[p].lo));
            const V inv = P::mk(R(1) / d.x, R(1) / d.y);
            row[p].lo = P::mul(row[p].lo, inv);
            row[p].hi = neg2_(P::mul(row[p].hi, inv));
          }
        }
      }
      site = nsite;
      thr = D::threshold(nthr_r);
      __syncwarp();  // the flip of this step is visible to the next one
    }

    // ---- sweep boundary: emit the configuration
    sync_sig();
    if (sweep >= a.n_discard && lane < W32)
      a.samples[(chain * a.chain_length + (sweep - a.n_discard)) * W32 + lane] = sig[lane];
  }
  sync_sig();
  if (lane < W32) a.chain_state[chain * W32 + lane] = sig[lane];
  if (a.n_accepted && lane == 0 && n_acc) atomicAdd(a.n_accepted, (unsigned long long)n_acc);
}

constexpr int PIPE_WR = PIPE_WR_;

template <typename R, int NP, int GL, int MB>
static int launch_ratio_pipe(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, const void *ws,
                             cudaStream_t st) {
  if (net->n_visible > 32 * PIPE_WR) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "pipelined sampler: N > %d", 32 * PIPE_WR);
  const int64_t blocks = (a.n_chains + SAMP_WARPS - 1) / SAMP_WARPS;
  const size_t smem = (size_t)SAMP_WARPS * ((size_t)PIPE_WR * 4 + 128 * (2 * sizeof(R) + 4));
  sample_ratio_pipe_kernel<R, NP, GL, MB, PIPE_WR><<<(unsigned)blocks, SAMP_WARPS * 32, smem, st>>>(
      make_netdev<R>(net), (const R *)ws, (R)a.s0, (R)a.s1, a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

// z <- z q on a pair of units with the results TIED to the registers of z (inline PTX, "+l" operands): with plain
// C++ the compiler keeps part of the new z in temporaries and reconciles the accept / multiply-back paths with a
// dozen register moves per chain and step.
__device__ __forceinline__ void cmul_inplace(float2 &zr, float2 &zi, float2 ql, float2 qh) {
  const float2 s = __fmul2_rn(zr, qh);
  const float2 t = __fmul2_rn(zi, qh);
  unsigned long long &ZR = reinterpret_cast<unsigned long long &>(zr), &ZI = reinterpret_cast<unsigned long long &>(zi);
  const unsigned long long QL = reinterpret_cast<const unsigned long long &>(ql);
  const float2 nt = make_float2(-t.x, -t.y);
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(ZI) : "l"(QL), "l"(reinterpret_cast<const unsigned long long &>(s)));
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(ZR) : "l"(QL), "l"(reinterpret_cast<const unsigned long long &>(nt)));
}
__device__ __forceinline__ void cmul_inplace(double2 &zr, double2 &zi, double2 ql, double2 qh) {
  const double2 s = make_double2(zr.x * qh.x, zr.y * qh.y), t = make_double2(zi.x * qh.x, zi.y * qh.y);
  zi = make_double2(fma(zi.x, ql.x, s.x), fma(zi.y, ql.y, s.y));
  zr = make_double2(fma(zr.x, ql.x, -t.x), fma(zr.y, ql.y, -t.y));
}

// ------------------------------------------------------------------------------------------------------------
// CTA-synchronous form for the shared site sequence (site_mode 0): the hot form of the multiplicative sampler.
//
// Every chain of the launch proposes the same site at the same step, so all chains of a CTA need the same two q
// rows (spin up / spin down) at step s.  The CTA stages them ONCE in shared memory, a chunk of T steps at a time,
// double-buffered with cp.async (global -> shared without registers) and one __syncthreads per chunk:
//   * L2 / L1 traffic per chain-step drops from one 16 NP-byte-per-lane row to 1 / (chains per CTA) of two rows
//     (the per-warp kernels re-read every row from L1 / L2: 4e11 B per launch at the north-star shape);
//   * no per-lane row registers: a warp carries C = 2..4 chains (z only, multiplied IN PLACE; a rejected move is
//     multiplied back by the other row of the same slot, which is the inverse), so 8..12 chains per SM
//     sub-partition are in flight instead of 4 and their reduction / decision latencies overlap;
//   * the site sequence and A_site come from a table written by prep_site_seq (they depend on the RNG alone).
// Random numbers -> steps exactly as in the other plain-target kernels.
constexpr int SITE_SEQ_CAP = 1 << 16;  // steps per call covered by the site table in the workspace

template <typename R>
__global__ void prep_site_seq(const R *A0, int N, uint32_t k0, uint32_t k1, uint64_t step_offset,
                              const unsigned long long *step_offset_dev, uint32_t n_steps, int *site_seq, R *a0_seq) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_steps) return;
  if (step_offset_dev) step_offset += *step_offset_dev;
  const uint64_t pos = step_offset + s, blk = pos >> 2;
  u32x4 c2 = {(uint32_t)blk, (uint32_t)(blk >> 32), 0u, (uint32_t)TAG_SITE_SHARED};
  const int site = (int)__umulhi(sel4(philox4x32_10(c2, k0, k1), (int)(pos & 3)), (uint32_t)N);
  site_seq[s] = site;
  a0_seq[s] = A0[site];
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <typename R, int NP, int C, int GL, int MAXW, int MINB>
__global__ void __launch_bounds__(32 * MAXW, MINB) sample_ratio_cta_kernel(NetDev<R> net, const R *ws, const R s0, const R s1,
                                                                   SampArgs a, const int *site_seq, const R *a0_seq,
                                                                   const int TSH) {
  using P = P2<R>;
  using V = typename P::V;
  using D = WarpDecide<R>;
  constexpr int L = 32, SB = 128;
  constexpr int NG = (NP + GL - 1) / GL;
  constexpr int ROW = 4 * NP * L;  // reals per (table, site) row
  extern __shared__ __align__(16) unsigned char cta_smem[];
  if (!samp_window_ok(a)) return;  // uniform over the grid: taken before any barrier
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int l = lane;
  const int T = 1 << TSH;  // steps per chunk
  R *rows = reinterpret_cast<R *>(cta_smem);                                   // [2][T][2][ROW]
  R *s_lu = rows + (size_t)4 * T * ROW + (size_t)warp * C * SB;                // [C][SB] of this warp
  // configurations of this warp's chains: [C][W32] words in shared memory (read with a broadcast LDS, flipped by
  // lane 0), which keeps 4 C registers free for a third chain per warp and lifts the N <= 128 limit
  uint32_t *sigx = reinterpret_cast<uint32_t *>(rows + (size_t)4 * T * ROW + (size_t)W * C * SB) + warp * C * W32;
  const int64_t chain0 = ((int64_t)blockIdx.x * W + warp) * C;
  const R *Qg = ws;  // global tables
  const V one = P::mk(R(1), R(1));
  // chains of this warp that exist (the others are carried along: every warp takes part in the CTA barriers)
  const int n_valid = a.n_chains - chain0 >= C ? C : (a.n_chains > chain0 ? (int)(a.n_chains - chain0) : 0);

  // ---- configurations
  for (int idx = lane; idx < C * W32; idx += 32) {
    const int c = idx / W32, w = idx - c * W32;
    uint32_t v = 0;
    if (c < n_valid) {
      if (a.init_random) {
        u32x4 ctr = {(uint32_t)(w >> 2), 0u, (uint32_t)(a.chain_offset + chain0 + c), (uint32_t)TAG_INIT};
        v = sel4(philox4x32_10(ctr, a.k0, a.k1), w & 3);
      } else {
        v = a.chain_state[(chain0 + c) * W32 + w];
      }
      const int nb = N - 32 * w;
      if (nb < 32) v &= (1u << nb) - 1u;
    }
    sigx[idx] = v;
  }
  __syncwarp();
  auto sync_sig = [&]() { __syncwarp(); };

  const int n_sweeps = a.n_discard + a.chain_length;
  const uint32_t n_steps = (uint32_t)n_sweeps * (uint32_t)a.sweep_size;
  const uint32_t phase = (uint32_t)(a.step_offset & (uint64_t)(SB - 1));
  const uint64_t batch0 = a.step_offset - phase;

  // log2 u of the batch that starts `pos0` steps after batch0, for the C chains of this warp
  auto refill = [&](uint32_t pos0) {
    __syncwarp();
    const uint64_t blk = ((batch0 + (uint64_t)pos0) >> 2) + (uint64_t)lane;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      u32x4 ctr = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)(a.chain_offset + chain0 + c), (uint32_t)TAG_UNIFORM};
      const u32x4 ru = philox4x32_10(ctr, a.k0, a.k1);
      s_lu[c * SB + 4 * lane + 0] = SMath<R>::log2u(ru.x);
      s_lu[c * SB + 4 * lane + 1] = SMath<R>::log2u(ru.y);
      s_lu[c * SB + 4 * lane + 2] = SMath<R>::log2u(ru.z);
      s_lu[c * SB + 4 * lane + 3] = SMath<R>::log2u(ru.w);
    }
    __syncwarp();
  };
  // request the 2 T rows of chunk qn (positions qn T .. qn T + T - 1) into buffer qn & 1
  auto issue_rows = [&](uint32_t qn) {
    for (int r = warp; r < 2 * T; r += W) {
      const int tt = r >> 1, tab = r & 1;
      const int64_t sidx = ((int64_t)qn << TSH) + tt - (int64_t)phase;
      const uint32_t sc = sidx < 0 ? 0u : (sidx >= (int64_t)n_steps ? n_steps - 1 : (uint32_t)sidx);
      const int st = __ldg(site_seq + sc);
      const char *src = reinterpret_cast<const char *>(Qg + (size_t)(tab * N + st) * ROW);
      char *dst = reinterpret_cast<char *>(rows + ((((qn & 1u) << TSH) + tt) * 2 + tab) * ROW);
      for (int off = lane * 16; off < (int)(ROW * sizeof(R)); off += 512) cp_async16(dst + off, src + off);
    }
  };

  V Zr[C][NP], Zi[C][NP];
  R lg[C];
  unsigned n_acc = 0;

  // z <- z * row (in place); returns this lane's sum_j log2 |1 + z_j|^2.  rp: shared memory, lane offset applied.
  // One group of GL pairs (one logarithm) at a time: stage by stage inside the group, which bounds the temporaries.
  auto multiply = [&](V (&zr)[NP], V (&zi)[NP], const R *rp, bool want_log) -> R {
    R acc = R(0);
#pragma unroll
    for (int gq = 0; gq < NG; ++gq) {
      constexpr int GS = GL;
      V qlo[GS], qhi[GS], m[GS];
#pragma unroll
      for (int i = 0; i < GS; ++i) {
        const int p = gq * GL + i;
        if (p < NP) {
          const V *q = reinterpret_cast<const V *>(rp + 4 * L * p);
          qlo[i] = q[0];
          qhi[i] = q[1];
        }
      }
#pragma unroll
      for (int i = 0; i < GS; ++i) {
        const int p = gq * GL + i;
        if (p < NP) cmul_inplace(zr[p], zi[p], qlo[i], qhi[i]);
      }
      if (want_log) {
#pragma unroll
        for (int i = 0; i < GS; ++i)
          if (gq * GL + i < NP) m[i] = add2_(zr[gq * GL + i], one);
#pragma unroll
        for (int i = 0; i < GS; ++i)
          if (gq * GL + i < NP) m[i] = P::mul(m[i], m[i]);
#pragma unroll
        for (int i = 0; i < GS; ++i)
          if (gq * GL + i < NP) m[i] = P::fma(zi[gq * GL + i], zi[gq * GL + i], m[i]);
#pragma unroll
        for (int st = 1; st < GS; st <<= 1)
#pragma unroll
          for (int i = 0; i < GS; ++i)
            if (i % (2 * st) == 0 && i + st < GS && gq * GL + i + st < NP) m[i] = P::mul(m[i], m[i + st]);
        const R lv = P::lg2(m[0].x * m[0].y);
        acc = gq == 0 ? lv : acc + lv;
      }
    }
    return acc;
  };

  // ---- prologue: rows of the first two chunks, random numbers of the first batch
  const uint32_t q0 = phase >> TSH;
  issue_rows(q0);
  cp_async_wait_all();
  __syncthreads();
  issue_rows(q0 + 1);
  refill(0);
  int nsite = __ldg(site_seq);
  R na0 = __ldg(a0_seq);

  uint32_t s = 0;  // steps done
  for (int sweep = 0; sweep < n_sweeps; ++sweep) {
    if ((sweep % RATIO_REFRESH) == 0) {
      // ---- theta from scratch (pairs of units), then z = exp(-2 theta); padded units carry z = 0
      sync_sig();
      const V *cW = reinterpret_cast<const V *>(net.W);
      const V *cb = reinterpret_cast<const V *>(net.b);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int j0 = 2 * L * p + l, j1 = j0 + L;
        const V b0 = (cb && j0 < M) ? cb[j0] : P::mk(R(0), R(0));
        const V b1 = (cb && j1 < M) ? cb[j1] : P::mk(R(0), R(0));
#pragma unroll
        for (int c = 0; c < C; ++c) {
          Zr[c][p] = P::mk(b0.x, b1.x);
          Zi[c][p] = P::mk(b0.y, b1.y);
        }
      }
      for (int i = 0; i < N; ++i) {
        V xx[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const R xv = ((sigx[c * W32 + (i >> 5)] >> (i & 31)) & 1u) ? s1 : s0;
          xx[c] = P::mk(xv, xv);
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int j0 = 2 * L * p + l, j1 = j0 + L;
          const V w0 = (j0 < M) ? __ldg(cW + (int64_t)i * M + j0) : P::mk(R(0), R(0));
          const V w1 = (j1 < M) ? __ldg(cW + (int64_t)i * M + j1) : P::mk(R(0), R(0));
#pragma unroll
          for (int c = 0; c < C; ++c) {
            Zr[c][p] = P::fma(xx[c], P::mk(w0.x, w1.x), Zr[c][p]);
            Zi[c][p] = P::fma(xx[c], P::mk(w0.y, w1.y), Zi[c][p]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int j0 = 2 * L * p + l, j1 = j0 + L;
          R e0 = exp_(R(-2) * Zr[c][p].x), e1 = exp_(R(-2) * Zr[c][p].y);
          R sn0, cs0, sn1, cs1;
          sincos_(R(-2) * Zi[c][p].x, &sn0, &cs0);
          sincos_(R(-2) * Zi[c][p].y, &sn1, &cs1);
          if (j0 >= M) e0 = R(0);
          if (j1 >= M) e1 = R(0);
          Zr[c][p] = P::mk(e0 * cs0, e1 * cs1);
          Zi[c][p] = P::mk(e0 * sn0, e1 * sn1);
        }
        R acc = R(0);
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
          V prod = one;
#pragma unroll
          for (int p = gq * GL; p < (gq + 1) * GL && p < NP; ++p) {
            const V aa = add2_(Zr[c][p], one);
            V m = P::mul(aa, aa);
            m = P::fma(Zi[c][p], Zi[c][p], m);
            prod = P::mul(prod, m);
          }
          acc += P::lg2(prod.x * prod.y);
        }
        lg[c] = acc;
      }
    }

    for (int k = 0; k < a.sweep_size; ++k, ++s) {
      const uint32_t pos = s + phase;
      const int tb = (int)(pos & (SB - 1));
      if (tb == 0 && s != 0) refill(pos);
      const int site = nsite;
      const R a0 = na0;
      {
        const uint32_t sn = s + 1 < n_steps ? s + 1 : s;
        nsite = __ldg(site_seq + sn);
        na0 = __ldg(a0_seq + sn);
      }
      const R *slot = rows + (pos & (uint32_t)(2 * T - 1)) * (2 * ROW) + 4 * l;  // [chunk parity][step in chunk]
      const uint32_t msk = 1u << (site & 31);
      const int wsel = site >> 5;

      uint32_t bit[C], word[C];
      R lgn[C];
#pragma unroll
      for (int c = 0; c < C; ++c) word[c] = sigx[c * W32 + wsel];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        bit[c] = (word[c] >> (site & 31)) & 1u;
        lgn[c] = multiply(Zr[c], Zi[c], slot + bit[c] * ROW, true);
        if (C > 2) asm volatile("" ::: "memory");  // keep the row loads of the next chain behind this one (registers)
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const typename D::T thr = D::threshold(s_lu[c * SB + tb] - (bit[c] ? -a0 : a0));
        if (D::accept(lgn[c] - lg[c], thr) && c < n_valid) {
          lg[c] = lgn[c];
          ++n_acc;
          if (lane == 0) sigx[c * W32 + wsel] = word[c] ^ msk;
        } else {
          // rejected: multiply back by the other row of the slot (the inverse of the one just applied)
          multiply(Zr[c], Zi[c], slot + (bit[c] ^ 1u) * ROW, false);
        }
      }
      __syncwarp();  // the flips of this step are visible to the next one
      if ((pos & (uint32_t)(T - 1)) == (uint32_t)(T - 1)) {
        // end of a chunk: the next one has landed everywhere, this one is free for the one after it
        cp_async_wait_all();
        __syncthreads();
        issue_rows((pos >> TSH) + 2);
      }
    }

    // ---- sweep boundary: emit the configurations
    sync_sig();
    if (sweep >= a.n_discard && lane < W32) {
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (c < n_valid)
          a.samples[((chain0 + c) * a.chain_length + (sweep - a.n_discard)) * W32 + lane] = sigx[c * W32 + lane];
    }
  }
  cp_async_wait_all();
  sync_sig();
  if (lane < W32) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (c < n_valid) a.chain_state[(chain0 + c) * W32 + lane] = sigx[c * W32 + lane];
  }
  if (a.n_accepted && lane == 0 && n_acc) atomicAdd(a.n_accepted, (unsigned long long)n_acc);
}

// warps per CTA and log2 of the chunk length: `maxw` warps (one CTA per SM) for large launches -- a wave of CTAs
// lasts (steps) x (latency of one step) almost independently of how many warps it holds, so fuller CTAs win over
// a better-filled last wave (measured); small launches get enough CTAs to touch every SM and shorter chunks.
inline void ratio_cta_shape(int64_t n_chains, int C, int sm_count, int maxw, int *W, int *TSH) {
  int64_t w = (n_chains + (int64_t)C * sm_count - 1) / ((int64_t)C * sm_count);
  if (w > maxw) w = maxw;
  if (w < 2) w = 2;
  *W = (int)w;
  *TSH = w >= 10 ? 3 : 2;
}
template <typename R>
inline size_t ratio_cta_smem(int NP, int C, int W, int TSH, int W32) {
  return ((size_t)4 * (1 << TSH) * 4 * NP * 32 + (size_t)W * C * 128) * sizeof(R) + (size_t)W * C * W32 * 4;
}

template <typename R, int NP, int C, int GL, int MAXW = 16, int MINB = 1>
static int launch_ratio_cta(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, const void *ws,
                            const int *site_seq, const R *a0_seq, cudaStream_t st, int force_w = 0) {
  int W, TSH;
  ratio_cta_shape(a.n_chains, C, h->sm_count, MAXW, &W, &TSH);
  if (force_w > 0) {  // the launch shares the GPU with others: fewer, full CTAs instead of one CTA on every SM
    W = force_w < MAXW ? force_w : MAXW;
    TSH = W >= 10 ? 3 : 2;
  }
  if (const char *e = getenv("NKFB_CTA_WARPS")) W = atoi(e) < MAXW ? atoi(e) : MAXW;  // tuning overrides
  if (const char *e = getenv("NKFB_CTA_TSH")) TSH = atoi(e);
  const size_t smem = ratio_cta_smem<R>(NP, C, W, TSH, (net->n_visible + 31) >> 5);
  cudaError_t rc = cudaFuncSetAttribute(sample_ratio_cta_kernel<R, NP, C, GL, MAXW, MINB>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (rc != cudaSuccess) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "sampler needs %zu B of shared memory", smem);
  const int64_t blocks = (a.n_chains + (int64_t)C * W - 1) / ((int64_t)C * W);
  sample_ratio_cta_kernel<R, NP, C, GL, MAXW, MINB><<<(unsigned)blocks, W * 32, smem, st>>>(
      make_netdev<R>(net), (const R *)ws, (R)a.s0, (R)a.s1, a, site_seq, a0_seq, TSH);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

template <typename R, int L, int NP, int C, int GL, int MB, int MODE = 1>
static int launch_ratio(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, const void *ws, cudaStream_t st) {
  constexpr int NCH = (32 / L) * C;
  const int W32 = (net->n_visible + 31) >> 5;
  const int64_t warps = (a.n_chains + NCH - 1) / NCH;
  const int64_t blocks = (warps + SAMP_WARPS - 1) / SAMP_WARPS;
  const size_t smem = (size_t)SAMP_WARPS * NCH * W32 * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    cudaError_t rc = cudaFuncSetAttribute(sample_ratio_kernel<R, L, NP, C, GL, MB, MODE>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "sampler needs %zu B of shared memory", smem);
  }
  sample_ratio_kernel<R, L, NP, C, GL, MB, MODE><<<(unsigned)blocks, SAMP_WARPS * 32, smem, st>>>(
      make_netdev<R>(net), (const R *)ws, (R)a.s0, (R)a.s1, a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
