// Shared device helpers for the nkfb200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges around the entry points (no-ops without a profiler attached)

#include "../../include/nkfb200.h"

struct nkfb_handle_s {
  int device;
  int sm_count;
  int64_t launches;
  char err[512];
  void *ws;        // scratch: gate matrix elements (nkfb_conn), sampler parameters when the caller gives none
  size_t ws_bytes;
  void *ws2;       // scratch: bf16 splits of the RBM kernels for the tensor-core estimator
  size_t ws2_bytes;
  void *ws3;       // scratch: bf16 splits of the RBM kernel for the tensor-core gradient
  size_t ws3_bytes;
  void *ws4;       // scratch: connected batch (packed rows, log-amplitudes, weights) of the Ising-type U^dagger gradient
  size_t ws4_bytes;
  int opt_sampler_algo;  // NKFB_OPT_SAMPLER_ALGO
  int opt_sampler_share; // NKFB_OPT_SAMPLER_SHARE: sampler launches of equal size that run side by side (default 1)
  int opt_fwd_max_ctas;  // NKFB_OPT_FWD_MAX_CTAS: upper bound on the CTAs of the tensor-core estimator (0 = none)
  int opt_tc_tables_valid;  // NKFB_OPT_TC_TABLES_VALID: the tables nkfb_tc_prepare wrote still match the parameters
  // what nkfb_tc_prepare (or the last estimator / gradient launch) left in ws2 / ws3: parameter buffers and geometry
  struct TcTablesKey {
    const void *w[2], *b[2];
    int n_visible, n_hidden, NN, NC, KC;
  } tc_fwd_key, tc_grad_key;
};

namespace nkfb {
inline bool tc_key_equal(const nkfb_handle_s::TcTablesKey &x, const nkfb_handle_s::TcTablesKey &y) {
  return x.w[0] == y.w[0] && x.w[1] == y.w[1] && x.b[0] == y.b[0] && x.b[1] == y.b[1] && x.n_visible == y.n_visible &&
         x.n_hidden == y.n_hidden && x.NN == y.NN && x.NC == y.NC && x.KC == y.KC && x.w[0] != nullptr;
}
}  // namespace nkfb

// Every entry point runs on the handle's device whatever the caller's current device is, and restores the latter.
struct NkfbDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit NkfbDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~NkfbDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define NKFB_DEVICE_GUARD(h) NkfbDeviceGuard nkfb_device_guard__((h)->device)

// NVTX range over one entry point (nsys / ncu --nvtx timelines: "nkfb_sample", "nkfb_infidelity_fwd", ...)
struct NkfbNvtxRange {
  explicit NkfbNvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NkfbNvtxRange() { nvtxRangePop(); }
};
#define NKFB_NVTX(name) NkfbNvtxRange nkfb_nvtx_range__(name)

#define NKFB_CHECK_CUDA(h, call)                                                        \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      snprintf((h)->err, sizeof((h)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, \
               cudaGetErrorString(e__));                                                \
      return NKFB_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define NKFB_FAIL(h, code, ...)                        \
  do {                                                 \
    snprintf((h)->err, sizeof((h)->err), __VA_ARGS__); \
    return (code);                                     \
  } while (0)

#define NKFB_LAUNCHED(h)                                                     \
  do {                                                                       \
    (h)->launches++;                                                         \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      snprintf((h)->err, sizeof((h)->err), "%s:%d: kernel launch failed: %s", \
               __FILE__, __LINE__, cudaGetErrorString(e__));                 \
      return NKFB_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

namespace nkfb {

constexpr double LN2 = 0.693147180559945309417232121458;
constexpr double LOG2E = 1.44269504088896340735992468100;

// ---------------------------------------------------------------- complex
template <typename R>
struct cx {
  R re, im;
};
template <typename R>
__host__ __device__ __forceinline__ cx<R> mk(R a, R b) {
  cx<R> z;
  z.re = a;
  z.im = b;
  return z;
}
template <typename R>
__device__ __forceinline__ cx<R> operator+(cx<R> a, cx<R> b) { return mk<R>(a.re + b.re, a.im + b.im); }
template <typename R>
__device__ __forceinline__ cx<R> operator-(cx<R> a, cx<R> b) { return mk<R>(a.re - b.re, a.im - b.im); }
template <typename R>
__device__ __forceinline__ cx<R> operator*(cx<R> a, cx<R> b) {
  return mk<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename R>
__device__ __forceinline__ cx<R> operator*(R a, cx<R> b) { return mk<R>(a * b.re, a * b.im); }
template <typename R>
__device__ __forceinline__ cx<R> conj(cx<R> a) { return mk<R>(a.re, -a.im); }
template <typename R>
__device__ __forceinline__ R abs2(cx<R> a) { return a.re * a.re + a.im * a.im; }

// ---------------------------------------------------------------- accurate math (estimator / gradient kernels)
__device__ __forceinline__ void sincos_(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float log_(float x) { return logf(x); }
__device__ __forceinline__ double log_(double x) { return log(x); }
__device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
__device__ __forceinline__ double log1p_(double x) { return log1p(x); }
__device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float hypot_(float y, float x) { return hypotf(y, x); }
__device__ __forceinline__ double hypot_(double y, double x) { return hypot(y, x); }

// log cosh z, NetKet's folding: sgn = -2 signbit(Re z) + 1; z *= sgn; z + log1p(exp(-2z)) - ln 2
// with the principal branch of the complex log1p.
template <typename R>
__device__ __forceinline__ cx<R> lncosh(cx<R> z) {
  R x = z.re, y = z.im;
  if (signbit(x)) {
    x = -x;
    y = -y;
  }
  R t = exp_(R(-2) * x);
  R s, c;
  sincos_(R(2) * y, &s, &c);
  R u = t * c, v = -t * s;
  R re = R(0.5) * log1p_(t * (R(2) * c + t));
  R im = atan2_(v, R(1) + u);
  return mk<R>(x + re - R(LN2), y + im);
}

// tanh z = sgn * ((1 - t^2) + 2 i t sin 2y) / (1 + 2 t cos 2y + t^2), t = exp(-2|x|)
template <typename R>
__device__ __forceinline__ cx<R> ctanh(cx<R> z) {
  R x = z.re, y = z.im;
  R sg = R(1);
  if (signbit(x)) {
    x = -x;
    y = -y;
    sg = R(-1);
  }
  R t = exp_(R(-2) * x);
  R s, c;
  sincos_(R(2) * y, &s, &c);
  R inv = sg / (R(1) + t * (R(2) * c + t));
  return mk<R>((R(1) - t * t) * inv, R(2) * t * s * inv);
}

template <typename R>
__device__ __forceinline__ cx<R> cexp(cx<R> z) {
  R e = exp_(z.re);
  R s, c;
  sincos_(z.im, &s, &c);
  return mk<R>(e * c, e * s);
}
template <typename R>
__device__ __forceinline__ cx<R> clog(cx<R> z) {
  return mk<R>(log_(hypot_(z.re, z.im)), atan2_(z.im, z.re));
}

// ---------------------------------------------------------------- fast MUFU math (sampler)
__device__ __forceinline__ float mufu_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_cos(float x) {
  float r;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_sin(float x) {
  float r;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
struct u32x4 {
  uint32_t x, y, z, w;
};
__host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    u32x4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// ---------------------------------------------------------------- bits
__device__ __forceinline__ int words_for(int N) { return (N + 31) >> 5; }

template <typename T>
__device__ __forceinline__ T shfl_xor_(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__device__ __forceinline__ double block_sum_256(double v, double *sh /* >= 8 doubles */) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

}  // namespace nkfb
