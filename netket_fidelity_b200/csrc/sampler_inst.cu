// Instantiations of the plain-target sampler for one (real type, lane-group width) pair.
// Compiled once per pair by the Makefile:  -DNKFB_RT=float|double -DNKFB_RN=f32|f64 -DNKFB_L=4|8|16|32
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

template <int K>
static int go(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, cudaStream_t st) {
  return launch_plain<NKFB_RT, NKFB_L, K, ChainsPerGroup<NKFB_RT, K>::value>(h, net, a, st);
}

int NKFB_FN(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int K, cudaStream_t st) {
  switch (K) {
    case 1: return go<1>(h, net, a, st);
    case 2: return go<2>(h, net, a, st);
    case 3: return go<3>(h, net, a, st);
    case 4: return go<4>(h, net, a, st);
    case 5: return go<5>(h, net, a, st);
    case 6: return go<6>(h, net, a, st);
    case 7: return go<7>(h, net, a, st);
    case 8: return go<8>(h, net, a, st);
    case 9: return go<9>(h, net, a, st);
    case 10: return go<10>(h, net, a, st);
    case 11: return go<11>(h, net, a, st);
    case 12: return go<12>(h, net, a, st);
    case 13: return go<13>(h, net, a, st);
    case 14: return go<14>(h, net, a, st);
    case 15: return go<15>(h, net, a, st);
    case 16: return go<16>(h, net, a, st);
#if NKFB_L == 32
    case 17: case 18: return go<18>(h, net, a, st);
    case 19: case 20: return go<20>(h, net, a, st);
    case 21: case 22: return go<22>(h, net, a, st);
    case 23: case 24: return go<24>(h, net, a, st);
    case 25: case 26: return go<26>(h, net, a, st);
    case 27: case 28: return go<28>(h, net, a, st);
    case 29: case 30: return go<30>(h, net, a, st);
    case 31: case 32: return go<32>(h, net, a, st);
#endif
    default: break;
  }
  NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "no sampler kernel for %d hidden units per lane", K);
}

}  // namespace nkfb
