// Handle lifecycle, (un)packing, materialised connected elements, chain statistics,
// micro-benchmarks.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"

namespace nkfb {

// ------------------------------------------------------------------ pack / unpack
template <typename R>
__global__ void pack_kernel(const R *x, int64_t B, int N, R s1, uint32_t *out) {
  const int W32 = (N + 31) >> 5;
  const int64_t total = B * W32;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / W32;
    const int w = (int)(idx - b * W32);
    uint32_t word = 0;
    const int i0 = w * 32, i1 = min(N, i0 + 32);
    for (int i = i0; i < i1; ++i)
      if (x[b * N + i] == s1) word |= (1u << (i - i0));
    out[idx] = word;
  }
}

template <typename R>
__global__ void unpack_kernel(const uint32_t *packed, int64_t B, int N, R s0, R s1, R *x) {
  const int W32 = (N + 31) >> 5;
  const int64_t total = B * N;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / N;
    const int i = (int)(idx - b * N);
    const uint32_t w = packed[b * W32 + (i >> 5)];
    x[idx] = ((w >> (i & 31)) & 1u) ? s1 : s0;
  }
}

// ------------------------------------------------------------------ materialised connected elements
// One thread per output element xp[b, c, i]; mels by the threads with i == 0.
template <typename R>
__global__ void conn_gate_kernel(const R *x, int64_t B, int N, int idx, R s0, R s1, const double *mel /*[2][2][2]*/,
                                 R *xp, double *mels) {
  const int64_t total = B * 2 * N;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % N);
    const int c = (int)((e / N) & 1);
    const int64_t b = e / (2 * (int64_t)N);
    const R v = x[b * N + i];
    R o = v;
    if (c == 1 && i == idx) o = (v == s0) ? s1 : s0;  // singlequbit_gates.py:96-99
    xp[e] = o;
    if (i == 0) {
      const int bit = (x[b * N + idx] == s0) ? 0 : 1;
      mels[2 * (b * 2 + c) + 0] = mel[(bit * 2 + c) * 2 + 0];
      mels[2 * (b * 2 + c) + 1] = mel[(bit * 2 + c) * 2 + 1];
    }
  }
}

template <typename R>
__global__ void conn_ising_kernel(const R *x, int64_t B, int N, R s0, R s1, const int32_t *edges, int n_edges,
                                  double c0r, double c0i, double c1r, double c1i, double c2r, double c2i, R *xp,
                                  double *mels) {
  const int C = N + 1;
  const int64_t total = B * (int64_t)C * N;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % N);
    const int64_t bc = e / N;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    const R v = x[b * N + i];
    R o = v;
    if (c >= 1 && i == c - 1) o = (v == s0) ? s1 : s0;
    xp[e] = o;
    if (i == 0) {
      if (c == 0) {
        double zz = 0;
        for (int q = 0; q < n_edges; ++q)
          zz += (x[b * N + edges[2 * q]] == x[b * N + edges[2 * q + 1]]) ? 1.0 : -1.0;
        mels[2 * bc + 0] = c0r + c1r * zz;
        mels[2 * bc + 1] = c0i + c1i * zz;
      } else {
        mels[2 * bc + 0] = c2r;
        mels[2 * bc + 1] = c2i;
      }
    }
  }
}

// ------------------------------------------------------------------ chain statistics partial sums
// One CTA per row. out[row*(3+nb) + {0: sum, 1: first half, 2: second half, 3+k: block k}].
template <typename R>
__global__ void chain_stats_kernel(const R *L, int64_t rows, int64_t cols, int64_t l_block, int64_t n_blocks,
                                   double *out) {
  __shared__ double red[8];
  const int64_t row = blockIdx.x;
  const R *p = L + row * cols;
  const int64_t half = cols / 2;
  double s = 0, h0 = 0, h1 = 0;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
    const double v = (double)p[c];
    s += v;
    if (c < half)
      h0 += v;
    else if (c < 2 * half)
      h1 += v;
  }
  double r;
  r = block_sum_256(s, red);
  if (threadIdx.x == 0) out[row * (3 + n_blocks) + 0] = r;
  r = block_sum_256(h0, red);
  if (threadIdx.x == 0) out[row * (3 + n_blocks) + 1] = r;
  r = block_sum_256(h1, red);
  if (threadIdx.x == 0) out[row * (3 + n_blocks) + 2] = r;
  // block sums: one warp per block, strided
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int64_t k = warp; k < n_blocks; k += nw) {
    double bs = 0;
    for (int64_t c = lane; c < l_block; c += 32) bs += (double)p[k * l_block + c];
    bs = warp_sum(bs);
    if (lane == 0) out[row * (3 + n_blocks) + 3 + k] = bs;
  }
}

// ------------------------------------------------------------------ micro-benchmarks
template <int KIND>
__global__ void microbench_kernel(int iters, float *sink) {
  float a = 0.001f * (threadIdx.x + 1), b = 0.5f + 0.001f * threadIdx.x, c = 1.0f + 0.01f * threadIdx.x,
        d = 0.25f + 0.002f * threadIdx.x;
  float e = a + 0.1f, f = b + 0.1f, g = c + 0.1f, hh = d + 0.1f;
  const float m1 = 0.999f + 1e-9f * (float)iters, m2 = 0.001f + 1e-9f * (float)iters;  // runtime values: no immediates
  float2 A2 = make_float2(a, e), B2 = make_float2(b, f), C2 = make_float2(c, g), D2 = make_float2(d, hh);
  double da = a, db = b, dc = c, dd = d, de = e, df = f, dg = g, dh = hh;
  const double dm1 = m1, dm2 = m2;
  for (int i = 0; i < iters; ++i) {
    if (KIND == 0) {
      a = mufu_ex2(a); b = mufu_ex2(b); c = mufu_ex2(c); d = mufu_ex2(d);
      e = mufu_ex2(e); f = mufu_ex2(f); g = mufu_ex2(g); hh = mufu_ex2(hh);
      a -= 1.f; b -= 1.f; c -= 1.f; d -= 1.f; e -= 1.f; f -= 1.f; g -= 1.f; hh -= 1.f;
    } else if (KIND == 1) {
      a = mufu_cos(a); b = mufu_cos(b); c = mufu_cos(c); d = mufu_cos(d);
      e = mufu_cos(e); f = mufu_cos(f); g = mufu_cos(g); hh = mufu_cos(hh);
    } else if (KIND == 2) {
      a = mufu_lg2(a); b = mufu_lg2(b); c = mufu_lg2(c); d = mufu_lg2(d);
      e = mufu_lg2(e); f = mufu_lg2(f); g = mufu_lg2(g); hh = mufu_lg2(hh);
      a = fabsf(a) + 1.f; b = fabsf(b) + 1.f; c = fabsf(c) + 1.f; d = fabsf(d) + 1.f;
      e = fabsf(e) + 1.f; f = fabsf(f) + 1.f; g = fabsf(g) + 1.f; hh = fabsf(hh) + 1.f;
    } else if (KIND == 3) {
      a = fmaf(a, 0.999f, 0.001f); b = fmaf(b, 0.999f, 0.001f); c = fmaf(c, 0.999f, 0.001f);
      d = fmaf(d, 0.999f, 0.001f); e = fmaf(e, 0.999f, 0.001f); f = fmaf(f, 0.999f, 0.001f);
      g = fmaf(g, 0.999f, 0.001f); hh = fmaf(hh, 0.999f, 0.001f);
    } else if (KIND == 4) {  // three-register FFMA
      a = fmaf(a, m1, m2); b = fmaf(b, m1, m2); c = fmaf(c, m1, m2); d = fmaf(d, m1, m2);
      e = fmaf(e, m1, m2); f = fmaf(f, m1, m2); g = fmaf(g, m1, m2); hh = fmaf(hh, m1, m2);
    } else if (KIND == 6) {  // FP64 FMA (the parity path's pipe): 8 independent DFMA chains kept in double
      da = fma(da, dm1, dm2); db = fma(db, dm1, dm2); dc = fma(dc, dm1, dm2); dd = fma(dd, dm1, dm2);
      de = fma(de, dm1, dm2); df = fma(df, dm1, dm2); dg = fma(dg, dm1, dm2); dh = fma(dh, dm1, dm2);
    } else {  // KIND 5: packed FFMA2 (2 FMAs per lane per instruction), 8 instructions = 16 FMAs
      float2 A = make_float2(a, b), B = make_float2(c, d), Cc = make_float2(e, f), D = make_float2(g, hh);
      const float2 M1 = make_float2(m1, m1), M2 = make_float2(m2, m2);
      A = __ffma2_rn(A, M1, M2); B = __ffma2_rn(B, M1, M2); Cc = __ffma2_rn(Cc, M1, M2); D = __ffma2_rn(D, M1, M2);
      A2 = __ffma2_rn(A2, M1, M2); B2 = __ffma2_rn(B2, M1, M2); C2 = __ffma2_rn(C2, M1, M2); D2 = __ffma2_rn(D2, M1, M2);
      a = A.x; b = A.y; c = B.x; d = B.y; e = Cc.x; f = Cc.y; g = D.x; hh = D.y;
    }
  }
  float r = a + b + c + d + e + f + g + hh + A2.x + A2.y + B2.x + B2.y + C2.x + C2.y + D2.x + D2.y;
  if (KIND == 6) r += (float)(da + db + dc + dd + de + df + dg + dh);
  if (r == 123.456f) sink[0] = r;
}

}  // namespace nkfb

using namespace nkfb;

extern "C" int nkfb_version(void) { return NKFB_VERSION; }

extern "C" int nkfb_create(int device, nkfb_handle_t *out) {
  if (!out) return NKFB_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return NKFB_ERR_CUDA;
  if (device < 0 || device >= count) return NKFB_ERR_INVALID;
  nkfb_handle_s *h = new (std::nothrow) nkfb_handle_s;
  if (!h) return NKFB_ERR_NOMEM;
  memset(h, 0, sizeof(*h));
  h->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete h;
    return NKFB_ERR_CUDA;
  }
  h->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    // sm_100a cubins only: refuse loudly instead of failing at first launch
    delete h;
    return NKFB_ERR_UNSUPPORTED;
  }
  {
    // initialise the context of `device` without leaving the caller's current device changed
    NkfbDeviceGuard guard(device);
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device || cudaFree(nullptr) != cudaSuccess) {
      delete h;
      return NKFB_ERR_CUDA;
    }
  }
  *out = h;
  return NKFB_OK;
}

extern "C" int nkfb_destroy(nkfb_handle_t h) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  if (h->ws) cudaFree(h->ws);
  if (h->ws2) cudaFree(h->ws2);
  if (h->ws3) cudaFree(h->ws3);
  if (h->ws4) cudaFree(h->ws4);
  delete h;
  return NKFB_OK;
}

extern "C" const char *nkfb_last_error(nkfb_handle_t h) { return h ? h->err : "null handle"; }

extern "C" int64_t nkfb_launch_count(nkfb_handle_t h) { return h ? h->launches : -1; }

static int grid_for(int64_t total, int sm) {
  int64_t g = (total + 255) / 256;
  int64_t cap = (int64_t)sm * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int nkfb_pack(nkfb_handle_t h, const void *x, int x_dtype, int64_t B, int N, const double ls[2],
                         uint32_t *packed, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  if (B <= 0) return NKFB_OK;
  if (!x || !packed || N < 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad pack arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int W32 = (N + 31) >> 5;
  int grid = grid_for(B * W32, h->sm_count);
  if (x_dtype == NKFB_F32)
    pack_kernel<float><<<grid, 256, 0, st>>>((const float *)x, B, N, (float)ls[1], packed);
  else if (x_dtype == NKFB_F64)
    pack_kernel<double><<<grid, 256, 0, st>>>((const double *)x, B, N, ls[1], packed);
  else
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad dtype");
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

extern "C" int nkfb_unpack(nkfb_handle_t h, const uint32_t *packed, int64_t B, int N, const double ls[2], void *x,
                           int x_dtype, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  if (B <= 0) return NKFB_OK;
  if (!x || !packed || N < 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad unpack arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(B * N, h->sm_count);
  if (x_dtype == NKFB_F32)
    unpack_kernel<float><<<grid, 256, 0, st>>>(packed, B, N, (float)ls[0], (float)ls[1], (float *)x);
  else if (x_dtype == NKFB_F64)
    unpack_kernel<double><<<grid, 256, 0, st>>>(packed, B, N, ls[0], ls[1], (double *)x);
  else
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad dtype");
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

extern "C" int nkfb_conn(nkfb_handle_t h, const nkfb_op_t *op, const void *x, int x_dtype, int64_t B, int N,
                         const double ls[2], void *xp, void *mels, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  if (!op || (op->kind != NKFB_OP_GATE && op->kind != NKFB_OP_ISING))
    NKFB_FAIL(h, NKFB_ERR_INVALID, "conn needs a GATE or ISING operator");
  if (B <= 0) return NKFB_OK;
  if (!x || !xp || !mels || N < 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad conn arguments");
  if (x_dtype != NKFB_F32 && x_dtype != NKFB_F64) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (op->kind == NKFB_OP_GATE) {
    if (op->idx < 0 || op->idx >= N) NKFB_FAIL(h, NKFB_ERR_INVALID, "gate site outside the lattice");
    // matrix elements travel through a small device buffer in the workspace
    if (h->ws_bytes < 64) {
      if (h->ws) cudaFree(h->ws);
      h->ws = nullptr;
      h->ws_bytes = 0;
      NKFB_CHECK_CUDA(h, cudaMalloc(&h->ws, 4096));
      h->ws_bytes = 4096;
    }
    NKFB_CHECK_CUDA(h, cudaMemcpyAsync(h->ws, op->gate_mel, 8 * sizeof(double), cudaMemcpyHostToDevice, st));
    int grid = grid_for(B * 2 * N, h->sm_count);
    if (x_dtype == NKFB_F32)
      conn_gate_kernel<float><<<grid, 256, 0, st>>>((const float *)x, B, N, op->idx, (float)ls[0], (float)ls[1],
                                                    (const double *)h->ws, (float *)xp, (double *)mels);
    else
      conn_gate_kernel<double><<<grid, 256, 0, st>>>((const double *)x, B, N, op->idx, ls[0], ls[1],
                                                     (const double *)h->ws, (double *)xp, (double *)mels);
  } else {
    if (op->n_edges > 0 && !op->edges) NKFB_FAIL(h, NKFB_ERR_INVALID, "ising operator without edges");
    int grid = grid_for(B * (int64_t)(N + 1) * N, h->sm_count);
    if (x_dtype == NKFB_F32)
      conn_ising_kernel<float><<<grid, 256, 0, st>>>((const float *)x, B, N, (float)ls[0], (float)ls[1], op->edges,
                                                     op->n_edges, op->c0[0], op->c0[1], op->c1[0], op->c1[1],
                                                     op->c2[0], op->c2[1], (float *)xp, (double *)mels);
    else
      conn_ising_kernel<double><<<grid, 256, 0, st>>>((const double *)x, B, N, ls[0], ls[1], op->edges, op->n_edges,
                                                      op->c0[0], op->c0[1], op->c1[0], op->c1[1], op->c2[0],
                                                      op->c2[1], (double *)xp, (double *)mels);
  }
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

extern "C" int nkfb_chain_stats(nkfb_handle_t h, const void *L, int dtype, int64_t rows, int64_t cols,
                                int64_t l_block, int64_t n_blocks, double *out, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_chain_stats");
  if (!L || !out || rows < 1 || cols < 1 || l_block < 1 || n_blocks < 0 || n_blocks * l_block > cols)
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad chain_stats arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NKFB_F32)
    chain_stats_kernel<float><<<(unsigned)rows, 256, 0, st>>>((const float *)L, rows, cols, l_block, n_blocks, out);
  else if (dtype == NKFB_F64)
    chain_stats_kernel<double><<<(unsigned)rows, 256, 0, st>>>((const double *)L, rows, cols, l_block, n_blocks,
                                                               out);
  else
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad dtype");
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

extern "C" int nkfb_set_option(nkfb_handle_t h, int key, int64_t value) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  switch (key) {
    case NKFB_OPT_SAMPLER_ALGO:
      if (value < 0 || value > 2) NKFB_FAIL(h, NKFB_ERR_INVALID, "sampler algorithm must be 0 (auto), 1 or 2");
      h->opt_sampler_algo = (int)value;
      return NKFB_OK;
    case NKFB_OPT_SAMPLER_SHARE:
      if (value < 1 || value > 16) NKFB_FAIL(h, NKFB_ERR_INVALID, "sampler share must be in 1..16");
      h->opt_sampler_share = (int)value;
      return NKFB_OK;
    case NKFB_OPT_FWD_MAX_CTAS:
      if (value < 0 || value > (1 << 20)) NKFB_FAIL(h, NKFB_ERR_INVALID, "estimator CTA limit must be >= 0");
      h->opt_fwd_max_ctas = (int)value;
      return NKFB_OK;
    case NKFB_OPT_TC_TABLES_VALID:
      if (value < 0 || value > 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "tables-valid flag must be 0 or 1");
      h->opt_tc_tables_valid = (int)value;
      return NKFB_OK;
    default: NKFB_FAIL(h, NKFB_ERR_INVALID, "unknown option %d", key);
  }
}

extern "C" int nkfb_microbench(nkfb_handle_t h, int kind, int iters, double *ops_per_second, void *stream) {
  if (!h || !ops_per_second || iters < 1) return NKFB_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  float *sink = nullptr;
  NKFB_CHECK_CUDA(h, cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  NKFB_CHECK_CUDA(h, cudaEventCreate(&e0));
  NKFB_CHECK_CUDA(h, cudaEventCreate(&e1));
  const int grid = h->sm_count * 8, block = 256;
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    NKFB_CHECK_CUDA(h, cudaEventRecord(e0, st));
    switch (kind) {
      case 0: microbench_kernel<0><<<grid, block, 0, st>>>(iters, sink); break;
      case 1: microbench_kernel<1><<<grid, block, 0, st>>>(iters, sink); break;
      case 2: microbench_kernel<2><<<grid, block, 0, st>>>(iters, sink); break;
      case 3: microbench_kernel<3><<<grid, block, 0, st>>>(iters, sink); break;
      case 4: microbench_kernel<4><<<grid, block, 0, st>>>(iters, sink); break;
      case 6: microbench_kernel<6><<<grid, block, 0, st>>>(iters, sink); break;
      default: microbench_kernel<5><<<grid, block, 0, st>>>(iters, sink); break;
    }
    h->launches++;
    NKFB_CHECK_CUDA(h, cudaEventRecord(e1, st));
    NKFB_CHECK_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    NKFB_CHECK_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)grid * block * (kind == 5 ? 16.0 : 8.0) * iters / (ms * 1e-3);
    if (rep > 0 && ops > best) best = ops;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *ops_per_second = best;
  return NKFB_OK;
}

// ------------------------------------------------------------------ optimiser update (driver loop)
namespace nkfb {
// kind 0: SGD  p -= lr g.   kind 1: Adam (optax.adam semantics; second moment uses |g|^2 for
// complex leaves): m = b1 m + (1-b1) g; v = b2 v + (1-b2)|g|^2; p -= lr * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)
template <typename R>
__global__ void optimizer_kernel(int kind, R *p, const R *g, R *m, R *v, int64_t P, double lr, double b1, double b2,
                                 double eps, double c1, double c2, int real_params) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    double gr = (double)g[2 * i], gi = real_params ? 0.0 : (double)g[2 * i + 1];
    if (kind == 0) {
      p[2 * i] = (R)((double)p[2 * i] - lr * gr);
      p[2 * i + 1] = (R)((double)p[2 * i + 1] - lr * gi);
    } else {
      double mr = b1 * (double)m[2 * i] + (1.0 - b1) * gr, mi = b1 * (double)m[2 * i + 1] + (1.0 - b1) * gi;
      double vv = b2 * (double)v[i] + (1.0 - b2) * (gr * gr + gi * gi);
      m[2 * i] = (R)mr;
      m[2 * i + 1] = (R)mi;
      v[i] = (R)vv;
      double den = sqrt(vv / c2) + eps;
      p[2 * i] = (R)((double)p[2 * i] - lr * (mr / c1) / den);
      p[2 * i + 1] = (R)((double)p[2 * i + 1] - lr * (mi / c1) / den);
    }
  }
}
}  // namespace nkfb

extern "C" int nkfb_optimizer_update(nkfb_handle_t h, int kind, int dtype, void *params, const void *grad, void *m,
                                     void *v, int64_t P, double lr, double b1, double b2, double eps, int64_t t,
                                     int real_params, void *stream) {
  if (!h) return NKFB_ERR_INVALID;
  NKFB_DEVICE_GUARD(h);
  NKFB_NVTX("nkfb_optimizer_update");
  if (!params || !grad || P <= 0 || (kind == 1 && (!m || !v)) || t < 1) NKFB_FAIL(h, NKFB_ERR_INVALID, "bad optimizer arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const double c1 = 1.0 - pow(b1, (double)t), c2 = 1.0 - pow(b2, (double)t);
  int grid = grid_for(P, h->sm_count);
  if (dtype == NKFB_F32)
    optimizer_kernel<float><<<grid, 256, 0, st>>>(kind, (float *)params, (const float *)grad, (float *)m, (float *)v, P,
                                                  lr, b1, b2, eps, c1, c2, real_params);
  else if (dtype == NKFB_F64)
    optimizer_kernel<double><<<grid, 256, 0, st>>>(kind, (double *)params, (const double *)grad, (double *)m,
                                                   (double *)v, P, lr, b1, b2, eps, c1, c2, real_params);
  else
    NKFB_FAIL(h, NKFB_ERR_INVALID, "bad dtype");
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}
