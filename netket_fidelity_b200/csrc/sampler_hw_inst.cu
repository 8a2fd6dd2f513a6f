// Instantiations of the half-warp lean sampler (sampler_lean_hw.cuh): single precision, shared site sequence, a chain
// on 16 lanes with NP16 pairs of hidden units per lane (M <= 32 NP16).  Own translation unit: the kernels are fully
// unrolled and compile in parallel with the 32-lane forms.
#include <cstdlib>

#include "sampler_lean_hw.cuh"

namespace nkfb {

bool ratio_hw_supported(int NP16) { return NP16 >= 11 && NP16 <= 16; }

template <int NP>
static int go_hw(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int gl_one, int warps, cudaStream_t st) {
#ifdef NKFB_TUNE
  if (const char *v = getenv("NKFB_HW_VARIANT")) {
    if (!gl_one) switch (atoi(v)) {
      case 1: return launch_ratio_lean_hw<NP, 4, 4, 16, 2>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 2: return launch_ratio_lean_hw<NP, 4, 2, 16, 2>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 3: return launch_ratio_lean_hw<NP, 4, 4, 16, 4>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 4: return launch_ratio_lean_hw<NP, 4, 4, 16, 1>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 5: return launch_ratio_lean_hw<NP, 4, 1, 16, 2>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 12: return launch_ratio_lean_hw<NP, 4, 1, 20, 1>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 13: return launch_ratio_lean_hw<NP, 4, 2, 20, 1>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 15: return launch_ratio_lean_hw<NP, 4, 8, 16, 1>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 16: return launch_ratio_lean_hw<NP, 4, 8, 16, 4>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 17: return launch_ratio_lean_hw<NP, 4, 8, 16, 3>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      case 14: return launch_ratio_lean_hw<NP, 4, 1, 16, 1>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
      default: break;
    }
  }
#endif
  // 16-warp CTAs only.  Measured on the 8-GPU shard of C4 (2048 chains; profiles/r02_sampler_notes.md, hw_check4/5.log):
  // 16 warps 5.15 ms; 14-warp CTAs of 28 chains (74 CTAs instead of 64) 6.66 ms; 16-warp CTAs of which 14 carry chains
  // (`warps` = 114) 6.62 ms; of which 12 carry chains (3 per sub-partition, `warps` = 112) 5.38 ms.  Fewer warps never
  // shorten the step and an uneven 4/4/3/3 split costs 29 %, so a shard below one wave keeps full CTAs on fewer SMs.
#ifdef NKFB_TUNE
  if (warps == 114) return launch_ratio_lean_hw<NP, 4, 4, 16, 2, 14>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  if (warps == 112) return launch_ratio_lean_hw<NP, 4, 4, 16, 2, 12>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  if (warps == 204) return launch_ratio_lean_hw<NP, 4, 8, 16, 1, 16, 1>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  if (warps == 216) return launch_ratio_lean_hw<NP, 4, 8, 16, 2, 16, 1>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  if (warps == 108) return launch_ratio_lean_hw<NP, 4, 4, 16, 2, 8>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  if (warps == 104) return launch_ratio_lean_hw<NP, 4, 4, 16, 2, 4>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  if (warps == 14) {
    if (gl_one) return launch_ratio_lean_hw<NP, 1, 4, 14>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
    return launch_ratio_lean_hw<NP, 4, 4, 14>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  }
#endif
  // tail of a large launch (the chains beyond its last full wave): 16-warp CTAs of which 8 warps carry chains -- 16 chains per
  // CTA, two warps per scheduler: a step lasts 0.69 of a full CTA's (3.54 against 5.12 ms per wave before the full unroll),
  // shorter than a round of the one-chain-per-warp kernel (3.92 ms)
  if (warps == 8) {
    if (gl_one) return launch_ratio_lean_hw<NP, 1, 8, 16, 2, 8>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
    return launch_ratio_lean_hw<NP, 4, 8, 16, 2, 8>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  }
  // (12 warps carrying chains -- 24 chains per CTA -- for a rest that does not fit 16 per CTA was measured too: the simulated
  // 4-GPU shard went from 11.62 to 12.74 ms; a wave of such CTAs lasts 4.87 ms with 43 of them but 5.38 ms with 86)
  // the 8 steps of a chunk fully unrolled (UNR = 8): 4.88 ms per wave against 5.13 (UNR = 4), 5.33 (2), 5.07 (1) at C4
  if (gl_one) return launch_ratio_lean_hw<NP, 1, 8>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
  return launch_ratio_lean_hw<NP, 4, 8>(h, net, a, a.ws_hw, a.site_seq, (const float *)a.a0_seq, st);
}

// gl_one = 0: one logarithm per 4 pairs;  gl_one = 1: one logarithm per pair (wider range of X)
int dispatch_ratio_hw_f32(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, int NP16, int gl_one, int warps,
                          cudaStream_t st) {
  switch (NP16) {
#ifndef NKFB_HW_ONLY13
    case 11: return go_hw<11>(h, net, a, gl_one, warps, st);
    case 12: return go_hw<12>(h, net, a, gl_one, warps, st);
    case 14: return go_hw<14>(h, net, a, gl_one, warps, st);
    case 15: return go_hw<15>(h, net, a, gl_one, warps, st);
    case 16: return go_hw<16>(h, net, a, gl_one, warps, st);
#endif
    case 13: return go_hw<13>(h, net, a, gl_one, warps, st);
    default: break;
  }
  NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "half-warp sampler: %d unit pairs per lane is not instantiated", NP16);
}

}  // namespace nkfb
