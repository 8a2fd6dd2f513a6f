// Instantiations of the plain-target sampler for one (real type, lane-group width) pair.
// Compiled once per pair by the Makefile:  -DNKFB_RT=float|double -DNKFB_RN=f32|f64 -DNKFB_L=4|8|16|32
#include <cstdlib>

#include "sampler_plain.cuh"

namespace nkfb {

#define NKFB_CAT_(a, b, c) a##b##_L##c
#define NKFB_CAT(a, b, c) NKFB_CAT_(a, b, c)
#define NKFB_FN NKFB_CAT(dispatch_plain_, NKFB_RN, NKFB_L)

// chains per lane group: bounded by the register file (2 K C registers of theta per lane)
template <typename R, int K>
struct ChainsPerGroup {
  static constexpr int value = sizeof(R) == 4 ? (K <= 8 ? 4 : (K <= 16 ? 2 : 1)) : (K <= 8 ? 2 : 1);
};

// minimum resident CTAs per SM (register cap): measured on B200 at K = 13 (north-star shape), two chains
// per group at <= 128 registers (4 CTAs = 16 warps / SM) beat the uncapped 176-register build by 17 %
template <typename R, int K>
struct MinBlocks {
  static constexpr int value = (sizeof(R) == 4 && K > 8 && K <= 16) ? 4 : 1;
};

template <int K>
static int go(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, const void *ws, cudaStream_t st) {
  return launch_plain<NKFB_RT, NKFB_L, K, ChainsPerGroup<NKFB_RT, K>::value, MinBlocks<NKFB_RT, K>::value>(h, net, a,
                                                                                                            ws, st);
}

int NKFB_FN(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int K, const void *ws, cudaStream_t st) {
#if NKFB_L == 32 && defined(NKFB_TUNE)
  // tuning variants of the north-star shape (M = 400 -> K = 13), selected by NKFB_SAMP_VARIANT
  if (K == 13 && sizeof(NKFB_RT) == 4) {
    const char *v = getenv("NKFB_SAMP_VARIANT");
    const int iv = v ? atoi(v) : 0;
    switch (iv) {
      case 1: return launch_plain<NKFB_RT, 32, 13, 2, 4>(h, net, a, ws, st);
      case 2: return launch_plain<NKFB_RT, 32, 13, 4, 1>(h, net, a, ws, st);
      case 3: return launch_plain<NKFB_RT, 32, 13, 4, 2>(h, net, a, ws, st);
      case 4: return launch_plain<NKFB_RT, 32, 13, 1, 4>(h, net, a, ws, st);
      case 5: return launch_plain<NKFB_RT, 32, 13, 1, 5>(h, net, a, ws, st);
      case 6: return launch_plain<NKFB_RT, 32, 13, 2, 3>(h, net, a, ws, st);
      case 7: return launch_plain<NKFB_RT, 32, 13, 1, 4, true>(h, net, a, ws, st);
      case 8: return launch_plain<NKFB_RT, 32, 13, 2, 4, true>(h, net, a, ws, st);
      case 9: return launch_plain<NKFB_RT, 32, 13, 1, 6, true>(h, net, a, ws, st);
      default: break;
    }
  }
#endif
#if NKFB_L == 16 && defined(NKFB_TUNE)
  if (K == 25 && sizeof(NKFB_RT) == 4) {
    const char *v = getenv("NKFB_SAMP_VARIANT");
    const int iv = v ? atoi(v) : 0;
    switch (iv) {
      case 11: return launch_plain<NKFB_RT, 16, 25, 1, 1, true>(h, net, a, ws, st);
      case 12: return launch_plain<NKFB_RT, 16, 25, 1, 4, true>(h, net, a, ws, st);
      case 13: return launch_plain<NKFB_RT, 16, 25, 1, 5, true>(h, net, a, ws, st);
      default: break;
    }
  }
#endif
#if NKFB_L == 8 && defined(NKFB_TUNE)
  if (K == 50 && sizeof(NKFB_RT) == 4) {
    const char *v = getenv("NKFB_SAMP_VARIANT");
    const int iv = v ? atoi(v) : 0;
    switch (iv) {
      case 21: return launch_plain<NKFB_RT, 8, 50, 1, 1, true>(h, net, a, ws, st);
      case 22: return launch_plain<NKFB_RT, 8, 50, 1, 3, true>(h, net, a, ws, st);
      default: break;
    }
  }
#endif
  switch (K) {
    case 1: return go<1>(h, net, a, ws, st);
    case 2: return go<2>(h, net, a, ws, st);
    case 3: return go<3>(h, net, a, ws, st);
    case 4: return go<4>(h, net, a, ws, st);
    case 5: return go<5>(h, net, a, ws, st);
    case 6: return go<6>(h, net, a, ws, st);
    case 7: return go<7>(h, net, a, ws, st);
    case 8: return go<8>(h, net, a, ws, st);
    case 9: return go<9>(h, net, a, ws, st);
    case 10: return go<10>(h, net, a, ws, st);
    case 11: return go<11>(h, net, a, ws, st);
    case 12: return go<12>(h, net, a, ws, st);
    case 13: return go<13>(h, net, a, ws, st);
    case 14: return go<14>(h, net, a, ws, st);
    case 15: return go<15>(h, net, a, ws, st);
    case 16: return go<16>(h, net, a, ws, st);
#if NKFB_L == 32
    case 17: return go<17>(h, net, a, ws, st);
    case 18: return go<18>(h, net, a, ws, st);
    case 19: return go<19>(h, net, a, ws, st);
    case 20: return go<20>(h, net, a, ws, st);
    case 21: return go<21>(h, net, a, ws, st);
    case 22: return go<22>(h, net, a, ws, st);
    case 23: return go<23>(h, net, a, ws, st);
    case 24: return go<24>(h, net, a, ws, st);
    case 25: return go<25>(h, net, a, ws, st);
    case 26: return go<26>(h, net, a, ws, st);
    case 27: return go<27>(h, net, a, ws, st);
    case 28: return go<28>(h, net, a, ws, st);
    case 29: return go<29>(h, net, a, ws, st);
    case 30: return go<30>(h, net, a, ws, st);
    case 31: return go<31>(h, net, a, ws, st);
    case 32: return go<32>(h, net, a, ws, st);
#endif
    default: break;
  }
  NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "no sampler kernel for %d hidden units per lane", K);
}

}  // namespace nkfb
