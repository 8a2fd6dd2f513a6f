// Shared definitions of the sampler kernels (see sampler.cu for the overview).
#pragma once
#include <cstdio>
#include <cstring>

#include "tile.cuh"

namespace nkfb {

enum { TAG_UNIFORM = 0, TAG_SITE_CHAIN = 1, TAG_SITE_SHARED = 2, TAG_INIT = 3 };

struct SampArgs {
  int64_t n_chains;
  int chain_length, n_discard, sweep_size, site_mode, init_random;
  uint32_t k0, k1;
  uint64_t chain_offset, step_offset;
  double s0, s1;
  uint32_t *chain_state, *samples;
  unsigned long long *n_accepted;
  // range window of the plain-target kernels (sampler_ratio.cuh): a kernel runs iff xbound == nullptr or
  // xlo < *xbound <= xhi, so that exactly one of the kernels launched back to back does the work
  const float *xbound;
  float xlo, xhi;
  // optional device-resident step counter added to step_offset (CUDA-graph replays: the host value is frozen)
  const unsigned long long *step_offset_dev;
  // site_mode 0: proposal sites / A_site of the steps of this call (prep_site_seq), or null
  const int *site_seq;
  const void *a0_seq;
  // q tables in the 16-lane layout for the half-warp kernel (sampler_lean_hw.cuh), np_hw pairs per lane; or null
  const void *ws_hw;
  int np_hw;
  int tables_valid;  // host side only: the workspace already holds this net's tables (NKFB_SAMPLE_TABLES_VALID)
  OpDev op;
};

__device__ __forceinline__ void samp_resolve_offset(SampArgs &a) {
  if (a.step_offset_dev) a.step_offset += *a.step_offset_dev;
}

__device__ __forceinline__ bool samp_window_ok(const SampArgs &a) {
  if (a.xbound == nullptr) return true;
  const float X = *a.xbound;
  return X > a.xlo && X <= a.xhi;
}

template <typename R>
struct SMath;
template <>
struct SMath<float> {
  // returns the log2-domain pieces: adds |x| to sumabs, multiplies prod by (1 + t^2 + 2 t cos 2y)
  static __device__ __forceinline__ void unit(float x, float y, float &sumabs, float &prod) {
    const float ax = fabsf(x);
    const float t = mufu_ex2(ax * (float)(-2.0 * LOG2E));
    const float c = mufu_cos(2.0f * y);
    prod *= fmaf(t, fmaf(2.0f, c, t), 1.0f);
    sumabs += ax;
  }
  static __device__ __forceinline__ float flush(float &prod) {
    float r = mufu_lg2(prod);
    prod = 1.0f;
    return r;
  }
  static __device__ __forceinline__ float finish(float sumabs, float lg) {
    return fmaf(sumabs, (float)LOG2E, 0.5f * lg);
  }
  static __device__ __forceinline__ float log2u(uint32_t r) {
    return mufu_lg2(((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f));
  }
};
template <>
struct SMath<double> {
  static __device__ __forceinline__ void unit(double x, double y, double &sumabs, double &prod) {
    const double ax = fabs(x);
    const double t = exp(-2.0 * ax);
    const double c = cos(2.0 * y);
    prod *= fma(t, fma(2.0, c, t), 1.0);
    sumabs += ax;
  }
  static __device__ __forceinline__ double flush(double &prod) {
    double r = log2(prod);
    prod = 1.0;
    return r;
  }
  static __device__ __forceinline__ double finish(double sumabs, double lg) { return fma(sumabs, LOG2E, 0.5 * lg); }
  static __device__ __forceinline__ double log2u(uint32_t r) { return log2(((double)r + 0.5) * (1.0 / 4294967296.0)); }
};

template <int L, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
  for (int m = L / 2; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__device__ __forceinline__ uint32_t sel4(const u32x4 &v, int c) {
  return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

constexpr int SAMP_WARPS = 4;


}  // namespace nkfb
