// Renyi-2 swap estimator (netket.experimental.observable.Renyi2EntanglementEntropy on MCState,
// exported at netket_fidelity/renyi2.py:1).
#include <cstdio>

#include "common.cuh"

namespace nkfb {

// first half sigma, second half sigma'.  out[p] = (sigma'_A, sigma_B), out[S+p] = (sigma_A, sigma'_B)
__global__ void renyi_swap_kernel(const uint32_t *samples, int64_t S_half, int W32, const uint32_t *mask,
                                  uint32_t *out) {
  const int64_t total = S_half * W32;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W32);
    const uint32_t m = mask[w];
    const uint32_t a = samples[idx], b = samples[total + idx];
    out[idx] = (b & m) | (a & ~m);
    out[total + idx] = (a & m) | (b & ~m);
  }
}

template <typename R>
__global__ void renyi_combine_kernel(const cx<R> *lp /* [4*S_half]: orig(2S) | swapped(2S) */, int64_t S_half,
                                     cx<R> *q, double *sums) {
  __shared__ double red[8];
  double sr = 0, sr2 = 0, si = 0, sn = 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < S_half;
       p += (int64_t)gridDim.x * blockDim.x) {
    const cx<R> l = lp[2 * S_half + p] + lp[3 * S_half + p] - lp[p] - lp[S_half + p];
    const cx<R> v = cexp(l);
    q[p] = v;
    sr += (double)v.re;
    sr2 += (double)v.re * (double)v.re;
    si += (double)v.im;
    sn += 1.0;
  }
  double v[4] = {sr, sr2, si, sn};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double r = block_sum_256(v[k], red);
    if (threadIdx.x == 0) atomicAdd(sums + k, r);
  }
}

}  // namespace nkfb

using namespace nkfb;

extern "C" int nkfb_renyi2(nkfb_handle_t h, const nkfb_rbm_t *net, const uint32_t *samples, int64_t S_half,
                           const uint32_t *mask, const double local_states[2], uint32_t *scratch,
                           void *scratch_logpsi, void *q, double *sums, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_renyi2");
  if (!net || !samples || !mask || !scratch || !scratch_logpsi || !q || !sums)
    NKFB_FAIL(h, NKFB_ERR_INVALID, "null buffer");
  if (S_half <= 0) return NKFB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int W32 = (net->n_visible + 31) >> 5;
  int64_t total = S_half * W32;
  int grid = (int)((total + 255) / 256 > (int64_t)h->sm_count * 16 ? (int64_t)h->sm_count * 16 : (total + 255) / 256);
  renyi_swap_kernel<<<grid, 256, 0, st>>>(samples, S_half, W32, mask, scratch);
  NKFB_LAUNCHED(h);
  const size_t esz = (net->dtype == NKFB_F32 ? 8 : 16);
  int rc = nkfb_logpsi(h, net, nullptr, samples, 2 * S_half, local_states, scratch_logpsi, stream);
  if (rc) return rc;
  rc = nkfb_logpsi(h, net, nullptr, scratch, 2 * S_half, local_states,
                   (char *)scratch_logpsi + esz * 2 * (size_t)S_half, stream);
  if (rc) return rc;
  int g2 = (int)((S_half + 255) / 256 > 1024 ? 1024 : (S_half + 255) / 256);
  if (net->dtype == NKFB_F32)
    renyi_combine_kernel<float><<<g2, 256, 0, st>>>((const cx<float> *)scratch_logpsi, S_half, (cx<float> *)q, sums);
  else
    renyi_combine_kernel<double><<<g2, 256, 0, st>>>((const cx<double> *)scratch_logpsi, S_half, (cx<double> *)q,
                                                     sums);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}
