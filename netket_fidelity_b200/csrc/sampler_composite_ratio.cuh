// Metropolis-local sampler for the composite target |sum_c mel_c phi(x'_c)|^2 of an ISING-type operator
// (mel_0 = c0 + c1 sum_<ij> z_i z_j on x itself, mel_{i+1} = c2 on x with site i flipped: IsingJax and the
// propagator 1 - i dt H of BASELINE config C3), MULTIPLICATIVE form.
//
// With z_j = exp(-2 theta_j(x)) and q_ij = exp(-2 d_i W_ij) (d_i = x_i^new - x_i^old) the amplitude of any
// configuration y reached from x by flipping a set S of sites is
//     phi(y) / phi(x) = prod_{i in S} E_i * prod_j (1 + z_j prod_{i in S} q_ij) / prod_j (1 + z_j),
//     E_i = exp(d_i (a_i + sum_j W_ij)),
// so, up to the common factor phi(x) / prod_j (1 + z_j) that cancels in the Metropolis ratio,
//     Phi(x)   ~ N(x)   = mdiag(x)  D + c2 sum_i E_i P(i),            D = prod_j (1 + z_j),
//     Phi(x^k) ~ N(x^k) = mdiag(x^k) E_k P(k) + c2 [E_k sum_{i != k} E_i P(k, i) + D],
//     P(S) = prod_j (1 + z_j prod_{i in S} q_ij).
// A proposal therefore costs N + 1 complex products over the M hidden units of u_j = z_j q_kj -- two complex
// multiplies and an add per (connected configuration, unit), in packed FP32 on pairs of units -- instead of N + 1
// complex log cosh sums (exp, sincos, log1p, atan2 per unit) and N + 1 warp reductions.  One warp per chain;
// lane r owns the connected configurations i = r, r + 32, ... (item N is the single flip, P(k)), whose products
// it accumulates with an integer exponent (renormalised every 8 pairs: no range restriction); z is recomputed
// from the configuration every sweep.  Random numbers -> steps exactly as in sample_composite_kernel.
#pragma once

#include "sampler_ratio.cuh"

namespace nkfb {

// ---- workspace (reals R):  Qc [2][MP][N][4]  |  E [2][N][4] = (Re E, Im E, log2 |E|^2, 0)
// (site innermost: the lanes of a warp own consecutive connected configurations i and read, for one pair of
// units, consecutive 16-byte entries -- with the site outermost every lane streams its own row and one LDG.128
// touches 32 cache lines)
inline size_t cr_q_reals(int N, int M) { return (size_t)8 * N * ((M + 1) / 2); }
__host__ __device__ inline size_t cr_q_reals_hd(int N, int M) { return (size_t)8 * N * ((M + 1) / 2); }
__host__ __device__ inline size_t cr_bound_off(int N, int M) { return cr_q_reals_hd(N, M) + (size_t)8 * N; }  // X (float), see sampler_ratio.cuh
inline size_t cr_ws_reals(int N, int M) { return cr_bound_off(N, M) + 8; }
// a run of 8 pairs between renormalisations multiplies 16 factors <= (1 + e^{2X}) / 2: finite in fp32 for X <= 3
constexpr float CR_XMAX = 3.0f;

template <typename R>
__global__ void prep_composite_ratio(NetDev<R> net, R *ws, R s0, R s1) {
  const int N = net.N, M = net.M, MP = (M + 1) / 2;
  const int64_t n_slots = (int64_t)2 * N * MP;
  const double dx = (double)s1 - (double)s0;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n_slots + 2 * N + M;
       idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx >= n_slots + 2 * N) {
      const int j = (int)(idx - n_slots - 2 * N);
      const double smax = fmax(fabs((double)s0), fabs((double)s1));
      double acc = 0.0;
      for (int i = 0; i < N; ++i) acc += fabs((double)net.W[2 * ((int64_t)i * M + j)]);
      acc = acc * smax + (net.b ? fabs((double)net.b[2 * j]) : 0.0);
      float xb = __double2float_ru(acc);
      if (!(xb == xb)) xb = __int_as_float(0x7f800000);
      atomicMax(reinterpret_cast<unsigned int *>(ws + cr_bound_off(N, M)), __float_as_uint(xb));
    } else if (idx < n_slots) {
      const int tab = idx >= (int64_t)N * MP;
      const int64_t rem = idx - (int64_t)tab * N * MP;
      const int i = (int)(rem / MP), p = (int)(rem - (int64_t)i * MP);
      const R sgn = (R)(tab == 0 ? -2.0 * dx : 2.0 * dx);
      R out[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = 2 * p + u;
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
      st4(ws + 4 * (((int64_t)tab * MP + p) * N + i), out);
    } else {
      const int64_t r = idx - n_slots;
      const int tab = r >= N, i = (int)(r - (int64_t)tab * N);
      double ar = net.a ? (double)net.a[2 * i] : 0.0, ai = net.a ? (double)net.a[2 * i + 1] : 0.0;
      for (int j = 0; j < M; ++j) {
        ar += (double)net.W[2 * ((int64_t)i * M + j)];
        ai += (double)net.W[2 * ((int64_t)i * M + j) + 1];
      }
      const double d = tab == 0 ? dx : -dx;
      const double e = exp(d * ar);
      R out[4] = {(R)(e * cos(d * ai)), (R)(e * sin(d * ai)), (R)(2.0 * LOG2E * d * ar), R(0)};
      st4(ws + 4 * (n_slots + r), out);
    }
  }
}

// value = (re + i im) 2^ex
template <typename R>
struct ScaledCx {
  R re, im;
  int ex;
};
__device__ __forceinline__ int ilogb_(float x) { return (int)((__float_as_uint(x) >> 23) & 255u) - 127; }
__device__ __forceinline__ int ilogb_(double x) { return (int)((__double2hiint(x) >> 20) & 2047) - 1023; }
__device__ __forceinline__ float pow2i_(int e, float) {
  e = e < -126 ? -126 : (e > 127 ? 127 : e);
  return __uint_as_float((uint32_t)(e + 127) << 23);
}
__device__ __forceinline__ double pow2i_(int e, double) {
  e = e < -1022 ? -1022 : (e > 1023 ? 1023 : e);
  return __hiloint2double((e + 1023) << 20, 0);
}
__device__ __forceinline__ float log2_(float x) { return log2f(x); }
__device__ __forceinline__ double log2_(double x) { return log2(x); }
template <typename R>
__device__ __forceinline__ void renorm(ScaledCx<R> &v) {
  const R m = fmax(fabs(v.re), fabs(v.im));
  if (m > R(0) && m < R(INFINITY)) {
    const int e = ilogb_(m);
    const R sc = pow2i_(-e, R(0));
    v.re *= sc;
    v.im *= sc;
    v.ex += e;
  }
}

template <typename R>
__global__ void __launch_bounds__(SAMP_WARPS * 32, sizeof(R) == 4 ? 7 : 3) sample_composite_ratio_kernel(NetDev<R> net, const R *ws,
                                                                                  SampArgs a) {
  using P = P2<R>;
  using V = typename P::V;
  using V4 = V4T<R>;
  extern __shared__ __align__(16) unsigned char cr_smem[];
  if (!samp_window_ok(a)) return;
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5, MP = (M + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t per_warp = ((size_t)8 * MP) * sizeof(R) + (((size_t)W32 * 4 + 15) & ~(size_t)15);
  unsigned char *base = cr_smem + (size_t)warp * per_warp;
  R *zs = reinterpret_cast<R *>(base);       // [MP][4] = (Re z_j0, Re z_j1, Im z_j0, Im z_j1)
  R *us = zs + (size_t)4 * MP;                // [MP][4]: u / 2 of the current proposal
  uint32_t *sg = reinterpret_cast<uint32_t *>(us + (size_t)4 * MP);
  const int64_t ch = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  if (ch >= a.n_chains) return;
  const R s0 = (R)a.s0, s1 = (R)a.s1;
  const OpDev &op = a.op;
  const R *Qc = ws;
  const R *Et = ws + cr_q_reals_hd(N, M);
  const cx<R> c0 = mk<R>((R)op.c0[0], (R)op.c0[1]), c1 = mk<R>((R)op.c1[0], (R)op.c1[1]);
  const cx<R> c2 = mk<R>((R)op.c2[0], (R)op.c2[1]);
  const V half2 = P::mk(R(0.5), R(0.5));

  for (int w = lane; w < W32; w += 32) {
    uint32_t v;
    if (a.init_random) {
      u32x4 ctr = {(uint32_t)(w >> 2), 0u, (uint32_t)(a.chain_offset + ch), (uint32_t)TAG_INIT};
      v = sel4(philox4x32_10(ctr, a.k0, a.k1), w & 3);
    } else {
      v = a.chain_state[ch * W32 + w];
    }
    const int nb = N - 32 * w;
    if (nb < 32) v &= (1u << nb) - 1u;
    sg[w] = v;
  }
  __syncwarp();
  auto spin_bit = [&](int i) -> uint32_t { return (sg[i >> 5] >> (i & 31)) & 1u; };

  // log2 |N|^2 of the configuration in `sg` flipped at site k (k < 0: the configuration itself), `us` = u / 2 with
  // u = z q_k (u = z for k < 0).  dzz: change of sum_<ij> z_i z_j under the flip.  Returns also the product of the
  // item N (the configuration without further flips) broadcast to every lane: the new D if the move is accepted.
  auto eval_N = [&](int k, R zz_new, const ScaledCx<R> &D, ScaledCx<R> &Pk_out) -> R {
    ScaledCx<R> term = {R(0), R(0), -(1 << 28)};
    cx<R> Ek = mk<R>(R(1), R(0));
    if (k >= 0) {
      const R *e = Et + 4 * ((size_t)(spin_bit(k)) * N + k);
      Ek = mk<R>(__ldg(e), __ldg(e + 1));
    }
    ScaledCx<R> mine_single = {R(0), R(0), 0};
    for (int it = lane; it <= N; it += 32) {
      ScaledCx<R> val;
      cx<R> coef;
      if (it == k) {  // flipping k twice: the configuration itself, amplitude D
        val = D;
        coef = c2;
      } else {
        V Pr = P::mk(R(1), R(1)), Pi = P::mk(R(0), R(0));
        int ex = 0;
        // one factor (1 + u q) / 2 (or (1 + u) / 2 for the item without a second flip) into the running products
        auto step_q = [&](const R *up, const R *qp) {
          const V4 u = *reinterpret_cast<const V4 *>(up);  // lo = Re pair, hi = Im pair
          const V4 q = V4::ld(qp);
          const V t = P::mul(u.hi, q.hi);
          const V vr = add2_(P::fma(u.lo, q.lo, neg2_(t)), half2);
          const V vi = P::fma(u.lo, q.hi, P::mul(u.hi, q.lo));
          const V t2 = P::mul(Pi, vi);
          const V nr = P::fma(Pr, vr, neg2_(t2));
          Pi = P::fma(Pr, vi, P::mul(Pi, vr));
          Pr = nr;
        };
        auto step_1 = [&](const R *up) {
          const V4 u = *reinterpret_cast<const V4 *>(up);
          const V vr = add2_(u.lo, half2);
          const V t2 = P::mul(Pi, u.hi);
          const V nr = P::fma(Pr, vr, neg2_(t2));
          Pi = P::fma(Pr, u.hi, P::mul(Pi, vr));
          Pr = nr;
        };
        auto renorm2 = [&]() {  // both running products by one common power of two
          const R m = fmax(fmax(fabs(Pr.x), fabs(Pr.y)), fmax(fabs(Pi.x), fabs(Pi.y)));
          if (m > R(0) && m < R(INFINITY)) {
            const int e = ilogb_(m);
            const R sc = pow2i_(-e, R(0));
            Pr = P::mul(Pr, P::mk(sc, sc));
            Pi = P::mul(Pi, P::mk(sc, sc));
            ex += 2 * e;
          }
        };
        const R *up = us;
        int p = 0;
        if (it < N) {
          const R *qp = Qc + 4 * ((size_t)spin_bit(it) * MP * N + it);
          const size_t qs = (size_t)4 * N;
          for (; p + 8 <= MP; p += 8) {
#pragma unroll
            for (int r = 0; r < 8; ++r) step_q(up + 4 * r, qp + qs * r);
            up += 32;
            qp += 8 * qs;
            renorm2();
          }
          for (; p < MP; ++p, up += 4, qp += qs) step_q(up, qp);
        } else {
          for (; p + 8 <= MP; p += 8) {
#pragma unroll
            for (int r = 0; r < 8; ++r) step_1(up + 4 * r);
            up += 32;
            renorm2();
          }
          for (; p < MP; ++p, up += 4) step_1(up);
        }
        val.re = Pr.x * Pr.y - Pi.x * Pi.y;
        val.im = Pr.x * Pi.y + Pi.x * Pr.y;
        val.ex = ex;
        renorm(val);
        if (it == N) {
          mine_single = val;
          const cx<R> md = mk<R>(c0.re + c1.re * zz_new, c0.im + c1.im * zz_new);
          coef = md * Ek;
        } else {
          const R *e = Et + 4 * ((size_t)(spin_bit(it)) * N + it);
          coef = c2 * (Ek * mk<R>(__ldg(e), __ldg(e + 1)));
        }
      }
      // term += coef * val (both scaled)
      ScaledCx<R> add = {coef.re * val.re - coef.im * val.im, coef.re * val.im + coef.im * val.re, val.ex};
      renorm(add);
      if (add.re != R(0) || add.im != R(0)) {
        if (add.ex >= term.ex) {
          const R sc = pow2i_(term.ex - add.ex, R(0));
          const bool tiny = term.ex - add.ex < -120;
          term.re = (tiny ? R(0) : term.re * sc) + add.re;
          term.im = (tiny ? R(0) : term.im * sc) + add.im;
          term.ex = add.ex;
        } else {
          const R sc = pow2i_(add.ex - term.ex, R(0));
          const bool tiny = add.ex - term.ex < -120;
          term.re += tiny ? R(0) : add.re * sc;
          term.im += tiny ? R(0) : add.im * sc;
        }
      }
    }
    // sum over the lanes at the largest exponent
    int emax = term.ex;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, m));
    const bool tiny = term.ex - emax < -120;
    const R sc = pow2i_(term.ex - emax, R(0));
    cx<R> sum = mk<R>(tiny ? R(0) : term.re * sc, tiny ? R(0) : term.im * sc);
    sum.re = warp_sum(sum.re);
    sum.im = warp_sum(sum.im);
    // the product of item N lives in lane N % 32
    const int src = N & 31;
    Pk_out.re = __shfl_sync(0xffffffffu, mine_single.re, src);
    Pk_out.im = __shfl_sync(0xffffffffu, mine_single.im, src);
    Pk_out.ex = __shfl_sync(0xffffffffu, mine_single.ex, src);
    return R(2) * (R)emax + log2_(sum.re * sum.re + sum.im * sum.im);  // |sum| is O(1): no range issue
  };

  unsigned n_acc = 0;
  const int total_sweeps = a.n_discard + a.chain_length;
  uint64_t step = a.step_offset;
  const uint32_t chid = (uint32_t)(a.chain_offset + ch);
  ScaledCx<R> D = {R(1), R(0), 0};
  R zz = R(0), lN = R(0);
  u32x4 rng_u = {0, 0, 0, 0}, rng_s = {0, 0, 0, 0};
  bool have_rng = false;

  for (int sweep = 0; sweep < total_sweeps; ++sweep) {
    // ---- refresh: z = exp(-2 theta) from the configuration, zz, D and log2 |N(x)|^2
    {
      __syncwarp();
      for (int p = lane; p < MP; p += 32) {
        R o[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = 2 * p + u;
          R zr = R(0), zi = R(0);
          if (j < M) {
            R tr = net.b ? net.b[2 * j] : R(0), ti = net.b ? net.b[2 * j + 1] : R(0);
            for (int i = 0; i < N; ++i) {
              const R x = spin_bit(i) ? s1 : s0;
              tr = fma(x, net.W[2 * ((int64_t)i * M + j)], tr);
              ti = fma(x, net.W[2 * ((int64_t)i * M + j) + 1], ti);
            }
            const R e = exp_(R(-2) * tr);
            R sn, cs;
            sincos_(R(-2) * ti, &sn, &cs);
            zr = e * cs;
            zi = e * sn;
          }
          o[u] = zr;
          o[2 + u] = zi;
        }
        st4(zs + 4 * p, o);
        R h[4] = {R(0.5) * o[0], R(0.5) * o[1], R(0.5) * o[2], R(0.5) * o[3]};
        st4(us + 4 * p, h);
      }
      R zl = R(0);
      for (int e = lane; e < op.n_edges; e += 32) {
        const int ia = op.edges[2 * e], ib = op.edges[2 * e + 1];
        zl += (spin_bit(ia) == spin_bit(ib)) ? R(1) : R(-1);
      }
      zz = warp_sum(zl);
      __syncwarp();
      ScaledCx<R> Pk;
      ScaledCx<R> dummy = {R(0), R(0), 0};
      lN = eval_N(-1, zz, dummy, Pk);  // k < 0: no item equals k, item N gives D = prod_j (1 + z_j) / 2
      // (eval_N uses D only for the item it == k, absent here; mdiag(x) multiplies the item-N product = D)
      D = Pk;
    }

    for (int t = 0; t < a.sweep_size; ++t, ++step) {
      const uint64_t blk = step >> 2;
      const int comp = (int)(step & 3);
      if (comp == 0 || !have_rng) {  // one Philox block serves four consecutive steps
        u32x4 ctr = {(uint32_t)blk, (uint32_t)(blk >> 32), chid, (uint32_t)TAG_UNIFORM};
        rng_u = philox4x32_10(ctr, a.k0, a.k1);
        if (a.site_mode == 0) {
          ctr.z = 0u;
          ctr.w = (uint32_t)TAG_SITE_SHARED;
        } else {
          ctr.w = (uint32_t)TAG_SITE_CHAIN;
        }
        rng_s = philox4x32_10(ctr, a.k0, a.k1);
        have_rng = true;
      }
      const uint32_t r_u = sel4(rng_u, comp);
      const uint32_t r_s = sel4(rng_s, comp);
      const int k = (int)__umulhi(r_s, (uint32_t)N);
      const R logu = SMath<R>::log2u(r_u);
      const uint32_t bk = spin_bit(k);

      // u / 2 = z q_k / 2
      {
        const R *qrow = Qc + 4 * ((size_t)bk * MP * N + k);
        for (int p = lane; p < MP; p += 32) {
          const V4 z = *reinterpret_cast<const V4 *>(zs + 4 * p);
          const V4 q = V4::ld(qrow + 4 * (size_t)p * N);
          const V t = P::mul(z.hi, q.hi);
          const V ur = P::mul(P::fma(z.lo, q.lo, neg2_(t)), half2);
          const V ui = P::mul(P::fma(z.lo, q.hi, P::mul(z.hi, q.lo)), half2);
          R o[4] = {ur.x, ur.y, ui.x, ui.y};
          st4(us + 4 * p, o);
        }
      }
      // change of sum_<ij> z_i z_j under the flip of k
      R dz = R(0);
      for (int e = lane; e < op.n_edges; e += 32) {
        const int ia = op.edges[2 * e], ib = op.edges[2 * e + 1];
        if (ia == k || ib == k) dz += (spin_bit(ia) == spin_bit(ib)) ? R(-2) : R(2);
      }
      dz = warp_sum(dz);
      __syncwarp();
      ScaledCx<R> Pk;
      const R lNn = eval_N(k, zz + dz, D, Pk);
      if (logu < lNn - lN) {
        // accepted: z <- u, D <- P(k), N <- N(x^k) / E_k
        const R l2E = __ldg(Et + 4 * ((size_t)bk * N + k) + 2);
        __syncwarp();
        for (int p = lane; p < MP; p += 32) {
          const V4 u = *reinterpret_cast<const V4 *>(us + 4 * p);
          R o[4] = {R(2) * u.lo.x, R(2) * u.lo.y, R(2) * u.hi.x, R(2) * u.hi.y};
          st4(zs + 4 * p, o);
        }
        if (lane == 0) {
          sg[k >> 5] ^= (1u << (k & 31));
          ++n_acc;
        }
        D = Pk;
        zz += dz;
        lN = lNn - l2E;
      }
      __syncwarp();
    }
    if (sweep >= a.n_discard) {
      const int ts = sweep - a.n_discard;
      for (int w = lane; w < W32; w += 32) a.samples[(ch * a.chain_length + ts) * W32 + w] = sg[w];
    }
  }
  for (int w = lane; w < W32; w += 32) a.chain_state[ch * W32 + w] = sg[w];
  if (a.n_accepted && lane == 0 && n_acc) atomicAdd(a.n_accepted, (unsigned long long)n_acc);
}

}  // namespace nkfb
