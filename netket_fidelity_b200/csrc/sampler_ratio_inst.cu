// Instantiations of the multiplicative ("ratio") plain-target sampler for one (real type, lane-group width)
// pair.  Compiled once per pair by the Makefile:  -DNKFB_RT=float|double -DNKFB_RN=f32|f64 -DNKFB_L=4|8|16|32
#include <algorithm>
#include <cstdlib>

#include "sampler_ratio.cuh"
#if NKFB_L == 32
#include "sampler_lean.cuh"
#include "sampler_lean_tma.cuh"
#endif

namespace nkfb {

// half-warp lean kernel (sampler_hw_inst.cu): a.ws_hw = q tables in the 16-lane layout, a.np_hw pairs per lane
int dispatch_ratio_hw_f32(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int NP16, int gl_one, int warps,
                          cudaStream_t st);

#define NKFB_CAT_(a, b, c) a##b##_L##c
#define NKFB_CAT(a, b, c) NKFB_CAT_(a, b, c)
#define NKFB_FN NKFB_CAT(dispatch_ratio_, NKFB_RN, NKFB_L)

// chains per lane group and minimum resident CTAs per SM: the state is 4 NP C registers (z) plus 4 NP for the row
// in flight, in units of sizeof(R) / 4
template <typename R, int NP>
struct RatioShape {
  static constexpr bool f32 = sizeof(R) == 4;
  static constexpr int C = f32 ? (NP <= 2 ? 4 : (NP <= 8 ? 2 : 1)) : (NP <= 2 ? 2 : 1);
  static constexpr int MB = f32 ? (NP <= 7 ? 4 : (NP <= 8 ? 3 : (NP <= 12 ? 4 : 3))) : (NP <= 4 ? 3 : 2);
};

// pipelined one-chain-per-warp kernel (32 lanes, N <= 128): three register images of 4 NP words (z, z', q row)
template <typename R, int NP>
struct PipeShape {
  static constexpr bool f32 = sizeof(R) == 4;
  static constexpr int MB = f32 ? (NP <= 7 ? 4 : (NP <= 9 ? 3 : 2)) : (NP <= 3 ? 4 : (NP <= 4 ? 3 : 2));
};

template <int NP>
static int go(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int gl_one, const void *ws, cudaStream_t st) {
  using S = RatioShape<NKFB_RT, NP>;
  constexpr int GW = NP < 4 ? NP : 4;
#if NKFB_L == 32
  if constexpr (NP <= 8) {
    // Which kernel runs how many chains (measured at the north-star shape, tools/shard_sim.py):
    //  * CTA-synchronous kernel, CC chains per warp, 16-warp CTAs, one per SM: a WAVE of CTAs lasts (steps) x
    //    (latency of one step) however full it is;
    //  * one-chain-per-warp pipelined kernel, 16 warps per SM: its step is about half as long, its time grows
    //    almost linearly with the chains beyond one round.
    // Both have the same throughput when full, so the choice is wave arithmetic: run k FULL waves on the
    // CTA-synchronous kernel and the remaining chains on the pipelined one, k chosen to minimise
    //    k x wave + max(1, rounds of the rest)          (wave = 1.70 rounds at CC = 2)
    // against ceil(waves) x wave for the CTA-synchronous kernel alone.  `share` launches of this size run side by
    // side (the two chain families of a step: NKFB_OPT_SAMPLER_SHARE = 2), each on sm_count / share SMs.  Same
    // Philox streams and site sequence everywhere: a chain's samples do not depend on the kernel up to
    // single-precision rounding of the log-ratio.
    constexpr int CC = sizeof(NKFB_RT) == 4 ? (NP <= 4 ? 4 : (NP <= 5 ? 3 : 2)) : 1;  // no register spills at 128
    const bool pipe_ok = net->n_visible <= 32 * PIPE_WR && !getenv("NKFB_RATIO_NOPIPE");
    const bool cta_ok = a.site_seq != nullptr && !getenv("NKFB_RATIO_NOCTA");
    int share = h->opt_sampler_share > 0 ? h->opt_sampler_share : 1;
    if (const char *e = getenv("NKFB_SAMPLER_SHARE")) share = atoi(e) > 0 ? atoi(e) : 1;
    if (share > 16) share = 16;
    const int64_t per_cta = (int64_t)CC * 16;
    const int64_t blocks = (a.n_chains + per_cta - 1) / per_cta;
    const int64_t slots = std::max<int64_t>(1, h->sm_count / share);
    const int64_t full_waves = blocks / slots;
    const double round_chains = 16.0 * h->sm_count / share;  // chains of this launch in one round of the pipelined kernel
    // head: chains on the CTA-synchronous kernel (0: none, n_chains: all); w16: its CTAs are 16 warps whatever their number
    int64_t head = 0;
    bool w16 = false;
    const bool legacy_big = a.n_chains >= (int64_t)40 * h->sm_count;
    if (cta_ok && (getenv("NKFB_RATIO_FORCECTA") || !pipe_ok)) {
      head = (legacy_big || getenv("NKFB_RATIO_FORCECTA") || getenv("NKFB_RATIO_NOPIPE")) ? a.n_chains : 0;
    } else if (cta_ok && getenv("NKFB_RATIO_HEADWAVES")) {  // tuning override: k full waves, the rest pipelined
      const int k = atoi(getenv("NKFB_RATIO_HEADWAVES"));
      head = k < 0 ? a.n_chains : std::min<int64_t>(a.n_chains, (int64_t)k * slots * per_cta);
      w16 = k >= 0;
    } else if (cta_ok && getenv("NKFB_RATIO_NOSPLIT")) {
      head = legacy_big ? a.n_chains : 0;
    } else if (cta_ok && CC == 2 && sizeof(NKFB_RT) == 4) {
      // one wave of the CTA-synchronous kernel in rounds of the pipelined one (3.92 ms per round at the C4 shape): 1.70 for
      // the round-1 kernel (6.65 ms per wave), 1.39 for the lean kernel with TMA row staging (5.46 ms)
      // 1.25 for the half-warp lean kernel (4.88 ms)
      const double wave = getenv("NKFB_NOLEAN") ? 1.70 : (a.ws_hw != nullptr ? 1.25 : 1.39);
      double best = (double)((blocks + slots - 1) / slots) * wave;  // CTA-synchronous kernel alone
      int64_t best_head = a.n_chains;
      // the rest: half-warp kernel with 8 of 16 warps carrying chains (16 chains per CTA, 0.87 of a pipelined round) when it
      // fits one wave of those, else rounds of the pipelined kernel
      const bool tail8 = a.ws_hw != nullptr && !getenv("NKFB_HW_NOTAIL8");
      for (int64_t k = full_waves; k >= 0; --k) {
        const int64_t hd = std::min<int64_t>(a.n_chains, k * slots * per_cta);
        const int64_t rest = a.n_chains - hd;
        const double t_rest = rest <= 0 ? 0.0 : ((tail8 && rest <= 16 * slots) ? 0.87 : std::max(1.0, (double)rest / round_chains));
        const double t = (double)k * wave + t_rest;
        if (t < best * 0.98) {
          best = t;
          best_head = hd;
        }
      }
      head = best_head;
      w16 = true;   // the model counts 16-warp CTAs: (sm_count / share) of them per wave and launch
    } else if (cta_ok && legacy_big) {
      // other shapes: all chains on the CTA-synchronous kernel, except a last partial wave that fits one round
      head = a.n_chains;
      int W, TSH, Wh = 0;
      ratio_cta_shape(a.n_chains, CC, h->sm_count, 16, &W, &TSH);
      const int64_t hd = full_waves * slots * per_cta;
      if (hd > 0) ratio_cta_shape(hd, CC, h->sm_count, 16, &Wh, &TSH);
      if (W == 16 && Wh == 16 && hd < a.n_chains && (double)(a.n_chains - hd) <= round_chains) head = hd;
    }
    if (head > 0) {
      SampArgs ah = a;
      ah.n_chains = head;
      const int fw = w16 ? 16 : 0;
      int rc;
      bool lean = false;
      if constexpr (sizeof(NKFB_RT) == 4 && CC == 2) lean = !getenv("NKFB_NOLEAN");
      if constexpr (sizeof(NKFB_RT) == 4 && CC == 2) {
        // single precision, two chains per warp: the lean form of the CTA-synchronous kernel (sampler_lean.cuh)
        // default: rows staged by the TMA unit with an mbarrier pipeline (sampler_lean_tma.cuh, 21.8 ms per launch at
        // C4); NKFB_LEAN_TMA=0: the cp.async form (sampler_lean.cuh, 22.6 ms)
        const char *tma = getenv("NKFB_LEAN_TMA");
        if (lean && a.ws_hw != nullptr) {
          // 14-warp CTAs for launches below one wave (tuning builds only: measured slower, sampler_hw_inst.cu)
          int warps = 16;
          if (getenv("NKFB_HW_W14") && head == a.n_chains && blocks < slots)
            warps = atoi(getenv("NKFB_HW_W14"));
          rc = dispatch_ratio_hw_f32(h, net, ah, a.np_hw, gl_one, warps, st);
        }
        else if (lean && !(tma && atoi(tma) == 0))
          rc = gl_one ? launch_ratio_lean_tma<NP, 2, 1, 4>(h, net, ah, ws, a.site_seq, (const float *)a.a0_seq, st)
                      : launch_ratio_lean_tma<NP, 2, GW, 4>(h, net, ah, ws, a.site_seq, (const float *)a.a0_seq, st);
        else if (lean)
          rc = gl_one ? launch_ratio_lean<NP, 2, 1>(h, net, ah, ws, a.site_seq, (const float *)a.a0_seq, st)
                      : launch_ratio_lean<NP, 2, GW>(h, net, ah, ws, a.site_seq, (const float *)a.a0_seq, st);
      }
      if (!lean)
        rc = gl_one ? launch_ratio_cta<NKFB_RT, NP, CC, 1>(h, net, ah, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st, fw)
                    : launch_ratio_cta<NKFB_RT, NP, CC, GW>(h, net, ah, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st, fw);
      if (rc || head == a.n_chains) return rc;
      const int W32 = (net->n_visible + 31) >> 5;
      SampArgs at = a;
      at.n_chains = a.n_chains - head;
      at.chain_offset = a.chain_offset + (uint64_t)head;
      at.chain_state = a.chain_state + head * W32;
      at.samples = a.samples ? a.samples + head * a.chain_length * W32 : nullptr;
      if constexpr (sizeof(NKFB_RT) == 4 && CC == 2) {
        if (lean && a.ws_hw != nullptr && at.n_chains <= 16 * slots && !getenv("NKFB_HW_NOTAIL8"))
          return dispatch_ratio_hw_f32(h, net, at, a.np_hw, gl_one, 8, st);
      }
      constexpr int PMB = PipeShape<NKFB_RT, NP>::MB;
      if (gl_one) return launch_ratio_pipe<NKFB_RT, NP, 1, PMB>(h, net, at, ws, st);
      return launch_ratio_pipe<NKFB_RT, NP, GW, PMB>(h, net, at, ws, st);
    }
  }
  if constexpr (sizeof(NKFB_RT) == 4 && NP >= 6 && NP <= 8) {
    // a whole launch of at most 16 chains per SM: the 8-of-16-warps form of the half-warp kernel (see the tail above)
    int share = h->opt_sampler_share > 0 ? h->opt_sampler_share : 1;
    if (a.ws_hw != nullptr && a.site_seq != nullptr && a.n_chains <= (int64_t)16 * std::max(1, h->sm_count / share) &&
        !getenv("NKFB_HW_NOTAIL8") && !getenv("NKFB_NOLEAN") && !getenv("NKFB_RATIO_NOCTA"))
      return dispatch_ratio_hw_f32(h, net, a, a.np_hw, gl_one, 8, st);
  }
  if (net->n_visible <= 32 * PIPE_WR && !getenv("NKFB_RATIO_NOPIPE")) {
    constexpr int PMB = PipeShape<NKFB_RT, NP>::MB;
    if (gl_one) return launch_ratio_pipe<NKFB_RT, NP, 1, PMB>(h, net, a, ws, st);
    return launch_ratio_pipe<NKFB_RT, NP, GW, PMB>(h, net, a, ws, st);
  }
#endif
  if (gl_one) return launch_ratio<NKFB_RT, NKFB_L, NP, S::C, 1, S::MB>(h, net, a, ws, st);
  return launch_ratio<NKFB_RT, NKFB_L, NP, S::C, GW, S::MB>(h, net, a, ws, st);
}

// NP = 9..16 (M up to 1024 on 32 lanes): single precision only
template <typename R>
static int go_big(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int NP, int gl_one, const void *ws,
                  cudaStream_t st) {
  if constexpr (sizeof(R) == 4) {
    switch (NP) {
      case 9: return go<9>(h, net, a, gl_one, ws, st);
      case 10: return go<10>(h, net, a, gl_one, ws, st);
      case 11: return go<11>(h, net, a, gl_one, ws, st);
      case 12: return go<12>(h, net, a, gl_one, ws, st);
      case 13: return go<13>(h, net, a, gl_one, ws, st);
      case 14: return go<14>(h, net, a, gl_one, ws, st);
      case 15: return go<15>(h, net, a, gl_one, ws, st);
      default: return go<16>(h, net, a, gl_one, ws, st);
    }
  }
  NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "ratio sampler: %d unit pairs per lane needs single precision", NP);
}

// gl_one = 0: one logarithm per min(NP, 4) pairs;  gl_one = 1: one logarithm per pair (wider range of X)
int NKFB_FN(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int NP, int gl_one, const void *ws,
            cudaStream_t st) {
#if NKFB_L == 32 && defined(NKFB_TUNE)
  // tuning variants of the north-star shape (M = 400 -> NP = 7), selected by NKFB_RATIO_VARIANT
  if (NP == 7 && sizeof(NKFB_RT) == 4 && !gl_one) {
    const char *v = getenv("NKFB_RATIO_VARIANT");
    const int iv = v ? atoi(v) : 0;
    switch (iv) {
      case 1: return launch_ratio<NKFB_RT, 32, 7, 2, 4, 3, 0>(h, net, a, ws, st);
      case 2: return launch_ratio<NKFB_RT, 32, 7, 2, 4, 4, 1>(h, net, a, ws, st);
      case 3: return launch_ratio<NKFB_RT, 32, 7, 1, 4, 4, 1>(h, net, a, ws, st);
      case 4: return launch_ratio<NKFB_RT, 32, 7, 1, 4, 5, 1>(h, net, a, ws, st);
      case 5: return launch_ratio<NKFB_RT, 32, 7, 1, 4, 6, 1>(h, net, a, ws, st);
      case 6: return launch_ratio<NKFB_RT, 32, 7, 2, 7, 4, 1>(h, net, a, ws, st);
      case 7: return launch_ratio<NKFB_RT, 32, 7, 2, 4, 3, 1>(h, net, a, ws, st);
      case 8: return launch_ratio<NKFB_RT, 32, 7, 3, 4, 3, 1>(h, net, a, ws, st);
      case 9: return launch_ratio<NKFB_RT, 32, 7, 1, 4, 5, 0>(h, net, a, ws, st);
      case 10: return launch_ratio<NKFB_RT, 32, 7, 2, 4, 5, 1>(h, net, a, ws, st);
      case 11: return launch_ratio<NKFB_RT, 32, 7, 4, 4, 2, 1>(h, net, a, ws, st);
      case 12: return launch_ratio<NKFB_RT, 32, 7, 3, 4, 4, 1>(h, net, a, ws, st);
      case 13: return launch_ratio<NKFB_RT, 32, 7, 1, 7, 6, 1>(h, net, a, ws, st);
      case 30: return launch_ratio_cta<NKFB_RT, 7, 3, 4>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 31: return launch_ratio_cta<NKFB_RT, 7, 2, 4>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 32: return launch_ratio_cta<NKFB_RT, 7, 4, 4>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 33: return launch_ratio_cta<NKFB_RT, 7, 3, 7>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 40: return launch_ratio_cta<NKFB_RT, 7, 1, 4, 14, 2>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 41: return launch_ratio_cta<NKFB_RT, 7, 1, 4, 10, 3>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 42: return launch_ratio_cta<NKFB_RT, 7, 1, 4, 16, 1>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 43: return launch_ratio_cta<NKFB_RT, 7, 1, 4, 8, 3>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 50: return launch_ratio_lean<7, 2, 4, 8, 16, 2>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 51: return launch_ratio_lean<7, 2, 4, 8, 16, 1>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 52: return launch_ratio_lean<7, 2, 4, 8, 16, 4>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 53: return launch_ratio_lean<7, 2, 4, 4, 16, 2>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 54: return launch_ratio_lean<7, 2, 7, 8, 16, 2>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 55: return launch_ratio_lean<7, 2, 7, 8, 16, 1>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 56: return launch_ratio_lean_tma<7, 2, 4, 8, 16, 2>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 57: return launch_ratio_lean_tma<7, 2, 4, 8, 16, 4>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 58: return launch_ratio_lean_tma<7, 2, 4, 4, 16, 2>(h, net, a, ws, a.site_seq, (const float *)a.a0_seq, st);
      case 44: return launch_ratio_cta<NKFB_RT, 7, 1, 4, 32, 1>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 45: return launch_ratio_cta<NKFB_RT, 7, 1, 2, 32, 1>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 46: return launch_ratio_cta<NKFB_RT, 7, 1, 4, 24, 1>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 47: return launch_ratio_cta<NKFB_RT, 7, 1, 4, 28, 1>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 34: return launch_ratio_cta<NKFB_RT, 7, 3, 4, 12>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 35: return launch_ratio_cta<NKFB_RT, 7, 4, 4, 8>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 36: return launch_ratio_cta<NKFB_RT, 7, 4, 4, 10>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 37: return launch_ratio_cta<NKFB_RT, 7, 3, 4, 14>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 38: return launch_ratio_cta<NKFB_RT, 7, 2, 4, 12>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 39: return launch_ratio_cta<NKFB_RT, 7, 5, 4, 8>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      case 20: return launch_ratio_pipe<NKFB_RT, 7, 4, 4>(h, net, a, ws, st);
      case 21: return launch_ratio_pipe<NKFB_RT, 7, 4, 3>(h, net, a, ws, st);
      case 22: return launch_ratio_pipe<NKFB_RT, 7, 4, 5>(h, net, a, ws, st);
      case 23: return launch_ratio_pipe<NKFB_RT, 7, 7, 4>(h, net, a, ws, st);
      case 24: return launch_ratio_pipe<NKFB_RT, 7, 2, 4>(h, net, a, ws, st);
      default: break;
    }
  }
#endif
  switch (NP) {
    case 1: return go<1>(h, net, a, gl_one, ws, st);
    case 2: return go<2>(h, net, a, gl_one, ws, st);
    case 3: return go<3>(h, net, a, gl_one, ws, st);
    case 4: return go<4>(h, net, a, gl_one, ws, st);
    case 5: return go<5>(h, net, a, gl_one, ws, st);
    case 6: return go<6>(h, net, a, gl_one, ws, st);
    case 7: return go<7>(h, net, a, gl_one, ws, st);
    case 8: return go<8>(h, net, a, gl_one, ws, st);
#if NKFB_L == 32
    case 9: case 10: case 11: case 12: case 13: case 14: case 15: case 16:
      return go_big<NKFB_RT>(h, net, a, NP, gl_one, ws, st);
#endif
    default: break;
  }
  NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "ratio sampler: %d unit pairs per lane is not instantiated", NP);
}

}  // namespace nkfb
