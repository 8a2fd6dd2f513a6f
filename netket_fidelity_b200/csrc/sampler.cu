// Metropolis-local sampler kernels.
//
// Plain RBM target (|psi|^2): a group of L lanes owns C Markov chains; lane l of the group
// holds the hidden pre-activations theta_j, j = l + L k (k < K <= KMAX), of each of its
// chains in registers and applies the incremental update theta += (x_new - x_old) W[i, :]
// for a single-site flip instead of a from-scratch forward.  The acceptance ratio needs
// only Re log cosh:  log2|cosh(x+iy)| = |x| log2(e) - 1 + 1/2 log2(1 + t^2 + 2 t cos 2y),
// t = 2^(-2 log2(e) |x|), i.e. MUFU.EX2 + MUFU.COS per hidden unit and one MUFU.LG2 per
// four units (the arguments of the logs are multiplied first).  Random numbers come from
// Philox4x32-10 keyed by the seed with (step block, global chain id, stream tag) as
// counter, so results do not depend on how chains are sharded over GPUs.
//
// Composite target (|sum_c mel_c phi(x'_c)|^2, c over the connected elements of a gate or
// of an Ising-type operator): one warp per chain, complex log cosh for every connected
// configuration of the proposal.
#include <algorithm>
#include <cstdlib>

#include "sampler_composite_ratio.cuh"

namespace nkfb {

// =====================================================================================
// Composite target: log Phi(x) = log sum_c mel_c(x) phi(x'_c).  One warp per chain.
template <typename R>
__device__ __forceinline__ cx<R> warp_csum(cx<R> v) {
  v.re = warp_sum(v.re);
  v.im = warp_sum(v.im);
  return v;
}

// log Phi of the configuration in `sg` whose pre-activations are (thr, thi) and a.x = ax
template <typename R, int KMAX>
__device__ __forceinline__ cx<R> composite_logamp(const NetDev<R> &net, const OpDev &op, const uint32_t *sg,
                                                  const R (&thr)[KMAX], const R (&thi)[KMAX], cx<R> ax, int K,
                                                  R s0, R s1) {
  const int N = net.N, M = net.M;
  const int lane = threadIdx.x & 31;
  const R ssum = s0 + s1;
  cx<R> base = mk<R>(0, 0);
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K && lane + 32 * k < M) base = base + lncosh(mk<R>(thr[k], thi[k]));
  base = warp_csum(base) + ax;
  if (op.kind == NKFB_OP_GATE) {
    const int idx = op.idx;
    const int bit = (sg[idx >> 5] >> (idx & 31)) & 1u;
    const R d = ssum - R(2) * (bit ? s1 : s0);
    cx<R> fl = mk<R>(0, 0);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int j = lane + 32 * k;
      if (k < K && j < M) {
        const R wr = net.W[2 * ((int64_t)idx * M + j)], wi = net.W[2 * ((int64_t)idx * M + j) + 1];
        fl = fl + lncosh(mk<R>(fma(d, wr, thr[k]), fma(d, wi, thi[k])));
      }
    }
    fl = warp_csum(fl) + ax;
    if (net.a) fl = fl + d * mk<R>(net.a[2 * idx], net.a[2 * idx + 1]);
    const cx<R> m0 = mk<R>((R)op.gate_mel[bit][0][0], (R)op.gate_mel[bit][0][1]);
    const cx<R> m1 = mk<R>((R)op.gate_mel[bit][1][0], (R)op.gate_mel[bit][1][1]);
    const R amax = fmax(base.re, fl.re);
    const cx<R> sum = m0 * cexp(mk<R>(base.re - amax, base.im)) + m1 * cexp(mk<R>(fl.re - amax, fl.im));
    cx<R> T = clog(sum);
    T.re += amax;
    return T;
  }
  // ISING: neighbours k = 0..N-1; lane (k % 32) keeps the amplitude of neighbour k in a small ring
  // (processed in rounds of 32 neighbours, online log-sum-exp across rounds)
  R zz = R(0);
  for (int e = lane; e < op.n_edges; e += 32) {
    const int ia = op.edges[2 * e], ib = op.edges[2 * e + 1];
    const int ba = (sg[ia >> 5] >> (ia & 31)) & 1u, bb = (sg[ib >> 5] >> (ib & 31)) & 1u;
    zz += (ba == bb) ? R(1) : R(-1);
  }
  zz = warp_sum(zz);
  const cx<R> mdiag = mk<R>((R)op.c0[0] + (R)op.c1[0] * zz, (R)op.c0[1] + (R)op.c1[1] * zz);
  const cx<R> moff = mk<R>((R)op.c2[0], (R)op.c2[1]);
  R run_max = base.re;
  cx<R> run_sum = mdiag * cexp(mk<R>(R(0), base.im));  // in units of exp(run_max)
  for (int k0 = 0; k0 < N; k0 += 32) {
    cx<R> mine = mk<R>(-INFINITY, 0);
    const int kend = min(32, N - k0);
    for (int kk = 0; kk < kend; ++kk) {
      const int k = k0 + kk;
      const int bit = (sg[k >> 5] >> (k & 31)) & 1u;
      const R d = ssum - R(2) * (bit ? s1 : s0);
      cx<R> fl = mk<R>(0, 0);
#pragma unroll
      for (int q = 0; q < KMAX; ++q) {
        const int j = lane + 32 * q;
        if (q < K && j < M) {
          const R wr = net.W[2 * ((int64_t)k * M + j)], wi = net.W[2 * ((int64_t)k * M + j) + 1];
          fl = fl + lncosh(mk<R>(fma(d, wr, thr[q]), fma(d, wi, thi[q])));
        }
      }
      fl = warp_csum(fl) + ax;
      if (net.a) fl = fl + d * mk<R>(net.a[2 * k], net.a[2 * k + 1]);
      if (lane == kk) mine = fl;
    }
    R mx = mine.re;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    const R new_max = fmax(run_max, mx);
    cx<R> part = mk<R>(0, 0);
    if (lane < kend) part = moff * cexp(mk<R>(mine.re - new_max, mine.im));
    part = warp_csum(part);
    const R sc = exp_(run_max - new_max);
    run_sum = mk<R>(run_sum.re * sc + part.re, run_sum.im * sc + part.im);
    run_max = new_max;
  }
  cx<R> T = clog(run_sum);
  T.re += run_max;
  return T;
}

template <typename R, int KMAX>
__global__ void __launch_bounds__(SAMP_WARPS * 32) sample_composite_kernel(NetDev<R> net, SampArgs a) {
  extern __shared__ uint32_t sig_all[];
  if (!samp_window_ok(a)) return;  // the multiplicative kernel covers this parameter range
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = (M + 31) / 32;
  uint32_t *sg = sig_all + (size_t)warp * W32;
  const int64_t ch = (int64_t)blockIdx.x * SAMP_WARPS + warp;
  if (ch >= a.n_chains) return;
  const R s0 = (R)a.s0, s1 = (R)a.s1, ssum = s0 + s1;
  const OpDev &op = a.op;  // read from the kernel arguments; no block-level sync is used

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

  R thr[KMAX], thi[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = lane + 32 * k;
    thr[k] = (k < K && j < M && net.b) ? net.b[2 * j] : R(0);
    thi[k] = (k < K && j < M && net.b) ? net.b[2 * j + 1] : R(0);
  }
  cx<R> ax = mk<R>(0, 0);
  for (int i = 0; i < N; ++i) {
    const R x = ((sg[i >> 5] >> (i & 31)) & 1u) ? s1 : s0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int j = lane + 32 * k;
      if (k < K && j < M) {
        thr[k] = fma(x, net.W[2 * ((int64_t)i * M + j)], thr[k]);
        thi[k] = fma(x, net.W[2 * ((int64_t)i * M + j) + 1], thi[k]);
      }
    }
    if (net.a) ax = ax + x * mk<R>(net.a[2 * i], net.a[2 * i + 1]);
  }
  R logp = R(2) * composite_logamp<R, KMAX>(net, op, sg, thr, thi, ax, K, s0, s1).re;

  unsigned n_acc = 0;
  const int total_sweeps = a.n_discard + a.chain_length;
  uint64_t step = a.step_offset;
  const uint32_t chid = (uint32_t)(a.chain_offset + ch);
  for (int sweep = 0; sweep < total_sweeps; ++sweep) {
    for (int t = 0; t < a.sweep_size; ++t, ++step) {
      const uint64_t blk = step >> 2;
      const int comp = (int)(step & 3);
      u32x4 ctr = {(uint32_t)blk, (uint32_t)(blk >> 32), chid, (uint32_t)TAG_UNIFORM};
      const uint32_t r_u = sel4(philox4x32_10(ctr, a.k0, a.k1), comp);
      if (a.site_mode == 0) {
        ctr.z = 0u;
        ctr.w = (uint32_t)TAG_SITE_SHARED;
      } else {
        ctr.w = (uint32_t)TAG_SITE_CHAIN;
      }
      const uint32_t r_s = sel4(philox4x32_10(ctr, a.k0, a.k1), comp);
      const int site = (int)__umulhi(r_s, (uint32_t)N);
      // natural-log uniform
      const R logu = SMath<R>::log2u(r_u) * (R)LN2;

      const uint32_t word = sg[site >> 5];
      const R xold = ((word >> (site & 31)) & 1u) ? s1 : s0;
      const R d = ssum - R(2) * xold;
      R pr[KMAX], pi[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int j = lane + 32 * k;
        pr[k] = thr[k];
        pi[k] = thi[k];
        if (k < K && j < M) {
          pr[k] = fma(d, net.W[2 * ((int64_t)site * M + j)], thr[k]);
          pi[k] = fma(d, net.W[2 * ((int64_t)site * M + j) + 1], thi[k]);
        }
      }
      cx<R> axn = ax;
      if (net.a) axn = axn + d * mk<R>(net.a[2 * site], net.a[2 * site + 1]);
      __syncwarp();
      if (lane == 0) sg[site >> 5] = word ^ (1u << (site & 31));  // tentatively flip
      __syncwarp();
      const R logpn = R(2) * composite_logamp<R, KMAX>(net, op, sg, pr, pi, axn, K, s0, s1).re;
      const bool acc = logu < (logpn - logp);
      if (acc) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          thr[k] = pr[k];
          thi[k] = pi[k];
        }
        ax = axn;
        logp = logpn;
        if (lane == 0) ++n_acc;
      } else {
        __syncwarp();
        if (lane == 0) sg[site >> 5] = word;  // undo
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

// ------------------------------------------------------------------ dispatch
// plain-target kernels are instantiated per (real type, lane-group width) in sampler_inst.cu
int dispatch_plain_f32_L4(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);
int dispatch_plain_f32_L8(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);
int dispatch_plain_f32_L16(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);
int dispatch_plain_f32_L32(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);
int dispatch_plain_f64_L4(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);
int dispatch_plain_f64_L8(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);
int dispatch_plain_f64_L16(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);
int dispatch_plain_f64_L32(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int K, const void *ws, cudaStream_t);

int dispatch_ratio_f32_L4(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);
int dispatch_ratio_f32_L8(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);
int dispatch_ratio_f32_L16(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);
int dispatch_ratio_f32_L32(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);
int dispatch_ratio_f64_L4(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);
int dispatch_ratio_f64_L8(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);
int dispatch_ratio_f64_L16(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);
int dispatch_ratio_f64_L32(nkfb_handle_t, const nkfb_rbm_t *, const SampArgs &, int NP, int gl_one, const void *ws, cudaStream_t);

// workspace = [ratio-sampler tables (sampler_ratio.cuh) | pre-scaled parameters of the additive sampler]
static size_t plain_ws_offset_reals(int N, int M) { return (ratio_ws_reals_max(N, M) + 63) & ~(size_t)63; }
static size_t seq_ws_offset_reals(int N, int M) { return (plain_ws_offset_reals(N, M) + sampler_ws_reals(N, M) + 63) & ~(size_t)63; }
// site table of the CTA-synchronous kernel: SITE_SEQ_CAP ints + SITE_SEQ_CAP reals
// ... followed by the q tables of the half-warp kernel in the 16-lane layout (sampler_lean_hw.cuh)
static size_t hw_ws_offset_reals(int N, int M) { return (seq_ws_offset_reals(N, M) + 2 * (size_t)SITE_SEQ_CAP + 63) & ~(size_t)63; }
static size_t sampler_ws_total_reals(int N, int M) { return hw_ws_offset_reals(N, M) + ratio_ws_reals_max(N, M); }
bool ratio_hw_supported(int NP16);

// lane-group width of the multiplicative sampler: least padded work among the instantiated shapes (<= 8 pairs per
// lane, <= 16 on 32 lanes in single precision); 0 if the shape is not covered
static int ratio_pick_L(int M, bool f32) {
  int bestL = 0, bestWork = 1 << 30;
  const int Ls[4] = {32, 16, 8, 4};
  for (int q = 0; q < 4; ++q) {
    const int L = Ls[q], NP = ratio_pairs(M, L);
    if (NP > (L == 32 && f32 ? 16 : 8)) continue;
    if (2 * L * NP < bestWork) {
      bestWork = 2 * L * NP;
      bestL = L;
    }
  }
  // the full-warp shapes have the pipelined / CTA-synchronous kernels: take them up to 30 % more padded work
  const int NP32 = ratio_pairs(M, 32);
  if (bestL != 0 && bestL != 32 && NP32 <= (f32 ? 16 : 8) && 10 * (64 * NP32) <= 13 * bestWork) bestL = 32;
  return bestL;
}

// Multiplicative sampler: launches the kernels whose range windows tile [0, x2]; returns the upper end x2 of
// the covered range in *covered (the additive kernel then takes X > x2), or leaves it at -1 if the shape is
// not covered at all.
static int dispatch_ratio(nkfb_handle_t h, const nkfb_rbm_t *net, SampArgs a, void *ws, cudaStream_t st,
                          float *covered) {
  const int M = net->n_hidden, N = net->n_visible;
  const bool f32 = net->dtype == NKFB_F32;
  const int bestL = ratio_pick_L(M, f32);
  if (bestL == 0) return NKFB_OK;  // not covered: the additive sampler does the whole range
  const int L = bestL, NP = ratio_pairs(M, L);
  const size_t a0_off = ratio_a0_off(N, M, L), b_off = ratio_bound_off(N, M, L);
  const size_t rsz = f32 ? 4 : 8;
  if (!a.tables_valid) {
    NKFB_CHECK_CUDA(h, cudaMemsetAsync((char *)ws + b_off * rsz, 0, rsz, st));
    const int64_t tot = (((int64_t)2 * N * NP * L + 31) & ~(int64_t)31) + (int64_t)(N + M) * 32;
    const int grid = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 8);
    if (f32)
      prep_ratio_params<float><<<grid, 256, 0, st>>>(make_netdev<float>(net), (float *)ws, L, NP, (float)a.s0,
                                                     (float)a.s1, a0_off, b_off);
    else
      prep_ratio_params<double><<<grid, 256, 0, st>>>(make_netdev<double>(net), (double *)ws, L, NP, a.s0, a.s1,
                                                      a0_off, b_off);
    NKFB_LAUNCHED(h);
  }
  a.xbound = reinterpret_cast<const float *>((const char *)ws + b_off * rsz);
  // shared site sequence on full warps: table of the proposal sites of this call for the CTA-synchronous kernel
  const uint64_t n_steps = (uint64_t)(a.n_discard + a.chain_length) * (uint64_t)a.sweep_size;
  if (a.site_mode == 0 && L == 32 && NP <= 8 && n_steps > 0 && n_steps <= (uint64_t)SITE_SEQ_CAP) {
    char *seq = (char *)ws + seq_ws_offset_reals(N, M) * rsz;
    int *site_seq = reinterpret_cast<int *>(seq);
    void *a0_seq = seq + (size_t)SITE_SEQ_CAP * 4;
    const unsigned grid = (unsigned)((n_steps + 255) / 256);
    if (f32)
      prep_site_seq<float><<<grid, 256, 0, st>>>((const float *)ws + a0_off, N, a.k0, a.k1, a.step_offset,
                                                 a.step_offset_dev, (uint32_t)n_steps, site_seq, (float *)a0_seq);
    else
      prep_site_seq<double><<<grid, 256, 0, st>>>((const double *)ws + a0_off, N, a.k0, a.k1, a.step_offset,
                                                  a.step_offset_dev, (uint32_t)n_steps, site_seq, (double *)a0_seq);
    NKFB_LAUNCHED(h);
    a.site_seq = site_seq;
    a.a0_seq = a0_seq;
    // single precision, two chains per warp (6 <= NP <= 8): the half-warp kernel takes the full CTAs of the launch; it
    // reads its own copy of the q tables in the 16-lane layout (the 32-lane copy serves the chains of the tail)
    const int NP16 = ratio_pairs(M, 16);
    if (f32 && NP >= 6 && ratio_hw_supported(NP16) && !getenv("NKFB_NOHW")) {
      float *ws16 = (float *)ws + hw_ws_offset_reals(N, M);
      if (!a.tables_valid) {
        const int64_t tot = (int64_t)2 * N * NP16 * 16;   // tables only: A_i and the range bound come from the 32-lane copy
        const int grid = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 8);
        prep_ratio_params<float><<<grid, 256, 0, st>>>(make_netdev<float>(net), ws16, 16, NP16, (float)a.s0, (float)a.s1,
                                                       (size_t)0, (size_t)0, /*tables_only=*/1);
        NKFB_LAUNCHED(h);
      }
      a.ws_hw = ws16;
      a.np_hw = NP16;
    }
  }
  const int GW = NP < 4 ? NP : 4;
  const float x1 = ratio_xmax(GW), x2 = ratio_xmax(1);
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && GW == 1) break;
    a.xlo = pass == 0 ? -1.0f : x1;
    a.xhi = pass == 0 ? x1 : x2;
    int rc;
    switch (L) {
      case 32: rc = f32 ? dispatch_ratio_f32_L32(h, net, a, NP, pass, ws, st) : dispatch_ratio_f64_L32(h, net, a, NP, pass, ws, st); break;
      case 16: rc = f32 ? dispatch_ratio_f32_L16(h, net, a, NP, pass, ws, st) : dispatch_ratio_f64_L16(h, net, a, NP, pass, ws, st); break;
      case 8: rc = f32 ? dispatch_ratio_f32_L8(h, net, a, NP, pass, ws, st) : dispatch_ratio_f64_L8(h, net, a, NP, pass, ws, st); break;
      default: rc = f32 ? dispatch_ratio_f32_L4(h, net, a, NP, pass, ws, st) : dispatch_ratio_f64_L4(h, net, a, NP, pass, ws, st); break;
    }
    if (rc != NKFB_OK) return rc;
  }
  *covered = x2;
  return NKFB_OK;
}

// lane-group width with the least padded work (ties -> wider group); K = ceil(M / L) <= 16 (32 for L = 32)
static int dispatch_plain(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a_in, void *ws, size_t ws_bytes,
                          cudaStream_t st) {
  const int M = net->n_hidden, N = net->n_visible;
  SampArgs a = a_in;
  // sampler tables in the caller's scratch or the handle's
  const size_t need = sampler_ws_total_reals(N, M) * (net->dtype == NKFB_F32 ? 4 : 8);
  if (ws) {
    if (ws_bytes < need) NKFB_FAIL(h, NKFB_ERR_INVALID, "sampler workspace too small: %zu < %zu bytes", ws_bytes, need);
  } else {
    if (h->ws_bytes < need) {
      if (h->ws) cudaFree(h->ws);
      h->ws = nullptr;
      h->ws_bytes = 0;
      NKFB_CHECK_CUDA(h, cudaMalloc(&h->ws, need));
      h->ws_bytes = need;
    }
    ws = h->ws;
  }
  // algorithm: 0 auto (multiplicative in single precision, additive in double), 1 additive, 2 multiplicative.
  // The multiplicative kernels cover X <= x2 (sampler_ratio.cuh); the additive kernel is always launched
  // behind them for X > x2, and alone when the shape is not covered.
  int algo = h->opt_sampler_algo;
  if (const char *e = getenv("NKFB_SAMPLER_ALGO")) algo = atoi(e);
  if (algo == 2 || (algo == 0 && net->dtype == NKFB_F32)) {
    float covered = -1.0f;
    const int rc = dispatch_ratio(h, net, a, ws, st, &covered);
    if (rc != NKFB_OK) return rc;
    if (covered >= 0.0f) {
      const int L = ratio_pick_L(M, net->dtype == NKFB_F32);
      a.xbound = reinterpret_cast<const float *>((const char *)ws +
                                                 ratio_bound_off(N, M, L) * (net->dtype == NKFB_F32 ? 4 : 8));
      a.xlo = covered;
      a.xhi = __builtin_inff();
    }
  }
  ws = (char *)ws + plain_ws_offset_reals(N, M) * (net->dtype == NKFB_F32 ? 4 : 8);
  if (!a.tables_valid) {
    const int64_t tot = (int64_t)N * M + M + N;
    const int grid = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 8);
    if (net->dtype == NKFB_F32)
      prep_sampler_params<float><<<grid, 256, 0, st>>>(make_netdev<float>(net), (float *)ws);
    else
      prep_sampler_params<double><<<grid, 256, 0, st>>>(make_netdev<double>(net), (double *)ws);
    NKFB_LAUNCHED(h);
  }
  int bestL = 0, bestWork = 1 << 30;
  const int Ls[4] = {32, 16, 8, 4};
  const char *forceL = getenv("NKFB_SAMP_L");  // tuning override
  for (int q = 0; q < 4; ++q) {
    const int L = Ls[q], K = (M + L - 1) / L;
    if (K > 16 && !(forceL && atoi(forceL) == L)) continue;
    if (forceL && atoi(forceL) == L) {
      bestL = L;
      break;
    }
    if (L * K < bestWork) {
      bestWork = L * K;
      bestL = L;
    }
  }
  if (bestL == 0) {
    if (M > 1024) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "n_hidden = %d > 1024 is not supported by the sampler", M);
    bestL = 32;
  }
  const int K = (M + bestL - 1) / bestL;
  const bool f32 = net->dtype == NKFB_F32;
  switch (bestL) {
    case 32: return f32 ? dispatch_plain_f32_L32(h, net, a, K, ws, st) : dispatch_plain_f64_L32(h, net, a, K, ws, st);
    case 16: return f32 ? dispatch_plain_f32_L16(h, net, a, K, ws, st) : dispatch_plain_f64_L16(h, net, a, K, ws, st);
    case 8: return f32 ? dispatch_plain_f32_L8(h, net, a, K, ws, st) : dispatch_plain_f64_L8(h, net, a, K, ws, st);
    default: return f32 ? dispatch_plain_f32_L4(h, net, a, K, ws, st) : dispatch_plain_f64_L4(h, net, a, K, ws, st);
  }
}

template <typename R, int KMAX>
static int launch_composite(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, cudaStream_t st) {
  const int W32 = (net->n_visible + 31) >> 5;
  const int64_t blocks = (a.n_chains + SAMP_WARPS - 1) / SAMP_WARPS;
  const size_t smem = (size_t)SAMP_WARPS * W32 * sizeof(uint32_t);
  sample_composite_kernel<R, KMAX><<<(unsigned)blocks, SAMP_WARPS * 32, smem, st>>>(make_netdev<R>(net), a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

// Ising-type operators: multiplicative kernel (sampler_composite_ratio.cuh) for X <= CR_XMAX, the log-cosh kernel
// behind it for the rest (same device-side range window as the plain-target kernels, no read-back)
template <typename R>
static int dispatch_composite_ratio(nkfb_handle_t h, const nkfb_rbm_t *net, SampArgs &a, void *ws, size_t ws_bytes,
                                    cudaStream_t st) {
  const int N = net->n_visible, M = net->n_hidden;
  const size_t need = cr_ws_reals(N, M) * sizeof(R);
  if (ws) {
    if (ws_bytes < need) NKFB_FAIL(h, NKFB_ERR_INVALID, "sampler workspace too small: %zu < %zu bytes", ws_bytes, need);
  } else {
    if (h->ws_bytes < need) {
      if (h->ws) cudaFree(h->ws);
      h->ws = nullptr;
      h->ws_bytes = 0;
      NKFB_CHECK_CUDA(h, cudaMalloc(&h->ws, need));
      h->ws_bytes = need;
    }
    ws = h->ws;
  }
  NKFB_CHECK_CUDA(h, cudaMemsetAsync((R *)ws + cr_bound_off(N, M), 0, sizeof(R), st));
  {
    const int64_t tot = (int64_t)2 * N * ((M + 1) / 2) + 2 * N + M;
    const int grid = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 8);
    prep_composite_ratio<R><<<grid, 256, 0, st>>>(make_netdev<R>(net), (R *)ws, (R)a.s0, (R)a.s1);
    NKFB_LAUNCHED(h);
  }
  a.xbound = reinterpret_cast<const float *>((const R *)ws + cr_bound_off(N, M));
  a.xlo = -1.0f;
  a.xhi = CR_XMAX;
  const int W32 = (N + 31) >> 5, MP = (M + 1) / 2;
  const size_t per_warp = (size_t)8 * MP * sizeof(R) + (((size_t)W32 * 4 + 15) & ~(size_t)15);
  const size_t smem = SAMP_WARPS * per_warp;
  if (smem > 200 * 1024) {
    a.xbound = nullptr;  // too many hidden units for the shared-memory state: log-cosh kernel only
    return NKFB_OK;
  }
  if (smem > 48 * 1024)
    NKFB_CHECK_CUDA(h, cudaFuncSetAttribute(sample_composite_ratio_kernel<R>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (a.n_chains + SAMP_WARPS - 1) / SAMP_WARPS;
  sample_composite_ratio_kernel<R><<<(unsigned)blocks, SAMP_WARPS * 32, smem, st>>>(make_netdev<R>(net), (const R *)ws, a);
  NKFB_LAUNCHED(h);
  a.xlo = CR_XMAX;
  a.xhi = __builtin_inff();
  return NKFB_OK;
}

template <typename R>
static int dispatch_composite(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, cudaStream_t st) {
  const int K = (net->n_hidden + 31) / 32;
  if (K <= 4) return launch_composite<R, 4>(h, net, a, st);
  if (K <= 8) return launch_composite<R, 8>(h, net, a, st);
  if (K <= 16) return launch_composite<R, 16>(h, net, a, st);
  NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "n_hidden = %d > 512 is not supported by the composite-target sampler",
            net->n_hidden);
}

}  // namespace nkfb

using namespace nkfb;

extern "C" int nkfb_sample(nkfb_handle_t h, const nkfb_rbm_t *net, const nkfb_op_t *op,
                           const nkfb_sample_cfg_t *cfg, uint32_t *chain_state, uint32_t *samples,
                           unsigned long long *n_accepted, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_sample");
  if (!net || !net->kernel || !cfg) NKFB_FAIL(h, NKFB_ERR_INVALID, "null argument");
  if (net->dtype != NKFB_F32 && net->dtype != NKFB_F64) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad RBM dtype");
  if (net->n_visible < 1 || net->n_hidden < 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad RBM shape");
  if (cfg->n_chains <= 0) return NKFB_OK;
  if (!chain_state || (!samples && cfg->chain_length > 0)) NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  if (cfg->chain_length < 0 || cfg->n_discard < 0 || cfg->sweep_size < 1)
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad chain_length / n_discard / sweep_size");
  if (cfg->chain_offset + (uint64_t)cfg->n_chains > 0xffffffffull)
    NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "more than 2^32 chains");
  if (op && op->kind == NKFB_OP_GATE && (op->idx < 0 || op->idx >= net->n_visible))
    NKFB_FAIL(h, NKFB_ERR_INVALID, "gate site outside the lattice");
  SampArgs a;
  memset(&a, 0, sizeof(a));
  a.n_chains = cfg->n_chains;
  a.chain_length = cfg->chain_length;
  a.n_discard = cfg->n_discard;
  a.sweep_size = cfg->sweep_size;
  a.site_mode = cfg->site_mode;
  a.init_random = cfg->init_random;
  a.k0 = (uint32_t)cfg->seed;
  a.k1 = (uint32_t)(cfg->seed >> 32);
  a.chain_offset = cfg->chain_offset;
  a.step_offset = cfg->step_offset;
  a.step_offset_dev = (const unsigned long long *)cfg->step_offset_dev;
  a.s0 = cfg->local_states[0];
  a.s1 = cfg->local_states[1];
  a.chain_state = chain_state;
  a.samples = samples;
  a.n_accepted = n_accepted;
  a.op = make_opdev(op);
  cudaStream_t st = (cudaStream_t)stream;
  const bool composite = op && op->kind != NKFB_OP_NONE;
  // tables of an earlier call: only with a caller-provided workspace (the handle's own buffer may have been reused)
  a.tables_valid = (cfg->flags & NKFB_SAMPLE_TABLES_VALID) && cfg->workspace != nullptr && !composite;
  if ((uint64_t)(cfg->n_discard + (int64_t)cfg->chain_length) * (uint64_t)cfg->sweep_size >= (1ull << 31))
    NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "more than 2^31 Metropolis steps in one call");
  if (!composite) return dispatch_plain(h, net, a, cfg->workspace, (size_t)cfg->workspace_bytes, st);
  int algo = h->opt_sampler_algo;
  if (const char *e = getenv("NKFB_SAMPLER_ALGO")) algo = atoi(e);
  if (op->kind == NKFB_OP_ISING && algo != 1) {
    const int rc = net->dtype == NKFB_F32
                       ? dispatch_composite_ratio<float>(h, net, a, cfg->workspace, (size_t)cfg->workspace_bytes, st)
                       : dispatch_composite_ratio<double>(h, net, a, cfg->workspace, (size_t)cfg->workspace_bytes, st);
    if (rc != NKFB_OK) return rc;
  }
  return net->dtype == NKFB_F32 ? dispatch_composite<float>(h, net, a, st) : dispatch_composite<double>(h, net, a, st);
}

extern "C" size_t nkfb_sample_workspace_bytes(int n_visible, int n_hidden, int dtype) {
  if (n_visible < 1 || n_hidden < 1) return 0;
  return sampler_ws_total_reals(n_visible, n_hidden) * (dtype == NKFB_F32 ? 4 : 8);
}
