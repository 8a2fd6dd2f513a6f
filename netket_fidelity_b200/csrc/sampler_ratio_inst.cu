// Instantiations of the multiplicative ("ratio") plain-target sampler for one (real type, lane-group width)
// pair.  Compiled once per pair by the Makefile:  -DNKFB_RT=float|double -DNKFB_RN=f32|f64 -DNKFB_L=4|8|16|32
#include <algorithm>
#include <cstdlib>

#include "sampler_ratio.cuh"

namespace nkfb {

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
    // small launches (a multi-GPU shard: fewer than ~40 chains per SM) are bound by the latency of one Metropolis
    // step, not by throughput: there the one-chain-per-warp pipelined kernel below, which has the shorter step, wins
    const bool big = a.n_chains >= (int64_t)40 * h->sm_count || getenv("NKFB_RATIO_NOPIPE") || getenv("NKFB_RATIO_FORCECTA");
    if (a.site_seq != nullptr && big && !getenv("NKFB_RATIO_NOCTA")) {
      // shared site sequence: CTA-synchronous kernel, C chains per warp (z only: 4 NP words per chain)
      constexpr int CC = sizeof(NKFB_RT) == 4 ? (NP <= 4 ? 4 : (NP <= 5 ? 3 : 2)) : 1;  // no register spills at 128
      // A wave of 16-warp CTAs (one per SM) lasts (steps) x (latency of one step) however full it is, so a last
      // wave that is less than half full costs as much as a full one.  When `share` launches of this size run side
      // by side (the two chain families of a step: NKFB_OPT_SAMPLER_SHARE = 2) and the CTAs beyond the last full
      // wave hold few enough chains for ONE round of the one-chain-per-warp kernel (8 warps per SM and launch at
      // 128 registers), those chains go to that kernel, whose step is about half as long.  Same Philox streams
      // and site sequence: a chain's samples do not depend on the kernel up to single-precision rounding.
      if (net->n_visible <= 32 * PIPE_WR && !getenv("NKFB_RATIO_NOSPLIT")) {
        int share = h->opt_sampler_share > 0 ? h->opt_sampler_share : 1;
        if (const char *e = getenv("NKFB_SAMPLER_SHARE")) share = atoi(e) > 0 ? atoi(e) : 1;
        int W, TSH;
        ratio_cta_shape(a.n_chains, CC, h->sm_count, 16, &W, &TSH);
        const int64_t per_cta = (int64_t)CC * W;
        const int64_t blocks = (a.n_chains + per_cta - 1) / per_cta;
        const int64_t slots = std::max<int64_t>(1, h->sm_count / share);
        const int64_t head = (blocks / slots) * slots * per_cta;
        const int64_t tail = a.n_chains - head;
        int Wh = 0;
        if (head > 0) ratio_cta_shape(head, CC, h->sm_count, 16, &Wh, &TSH);
        if (W == 16 && Wh == 16 && tail > 0 && tail <= (int64_t)(16 / (share < 16 ? share : 16)) * h->sm_count) {
          const int W32 = (net->n_visible + 31) >> 5;
          SampArgs ah = a, at = a;
          ah.n_chains = head;
          at.n_chains = tail;
          at.chain_offset = a.chain_offset + (uint64_t)head;
          at.chain_state = a.chain_state + head * W32;
          at.samples = a.samples ? a.samples + head * a.chain_length * W32 : nullptr;
          int rc = gl_one ? launch_ratio_cta<NKFB_RT, NP, CC, 1>(h, net, ah, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st)
                          : launch_ratio_cta<NKFB_RT, NP, CC, GW>(h, net, ah, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
          if (rc) return rc;
          constexpr int PMB = PipeShape<NKFB_RT, NP>::MB;
          if (gl_one) return launch_ratio_pipe<NKFB_RT, NP, 1, PMB>(h, net, at, ws, st);
          return launch_ratio_pipe<NKFB_RT, NP, GW, PMB>(h, net, at, ws, st);
        }
      }
      if (gl_one)
        return launch_ratio_cta<NKFB_RT, NP, CC, 1>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
      return launch_ratio_cta<NKFB_RT, NP, CC, GW>(h, net, a, ws, a.site_seq, (const NKFB_RT *)a.a0_seq, st);
    }
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
