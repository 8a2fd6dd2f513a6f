// Microbenchmarks behind the sampler design: (1) fma.rn.f32x2 issue rate of ONE warp as a function of the number of
// independent accumulators (ILP) and of the warps per SM sub-partition; (2) the dense part of one Metropolis step of
// the multiplicative sampler (7 pairs: complex multiply by a row from shared memory, |1+z|^2, product tree, lg2)
// without any decision, per warp-step, as a function of warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pk(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ float lo(u64 v) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return x; }
__device__ __forceinline__ float hi(u64 v) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return y; }

template <int ILP, bool PACKED>
__global__ void k_fma(float *out, int iters, float s) {
  u64 acc[ILP]; float fa[ILP];
  const u64 b = pk(s, s * 1.0001f), c = pk(1e-9f, 2e-9f);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { acc[i] = pk(threadIdx.x * 1e-3f + i, i * 0.5f); fa[i] = threadIdx.x * 1e-3f + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (PACKED) acc[i] = fma2(acc[i], b, c); else fa[i] = fmaf(fa[i], s, 1e-9f);
      }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += PACKED ? lo(acc[i]) + hi(acc[i]) : fa[i];
  if (r == 123.456f) out[0] = r;
}

// dense part of a step: C chains x NP pairs; rows in shared memory [2][NP][32] float4
template <int C, int NP, int MODE>
__global__ void k_dense(float *out, int iters) {
  extern __shared__ float4 rows[];
  for (int i = threadIdx.x; i < 2 * NP * 32; i += blockDim.x) rows[i] = make_float4(1.0001f, 0.9999f, 1e-4f, -1e-4f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  u64 zr[C][NP], zi[C][NP];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int p = 0; p < NP; ++p) { zr[c][p] = pk(0.1f + lane * 1e-3f, 0.2f); zi[c][p] = pk(0.05f, 0.01f * p); }
  const u64 one = pk(1.f, 1.f);
  float acc = 0;
  unsigned bit = lane & 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 *rp = rows + ((bit ^ c) & 1) * NP * 32 + lane;
      u64 m[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        float4 q;
        if (MODE == 1) q = make_float4(1.0001f + 1e-6f * p, 0.9999f, 1e-4f, -1e-4f + 1e-7f * it); else q = rp[p * 32];
        const u64 ql = pk(q.x, q.y), qh = pk(q.z, q.w);
        if (MODE == 2) { zr[c][p] ^= ql; zi[c][p] ^= qh; continue; }
        const u64 s = mul2(zr[c][p], qh), t = mul2(zi[c][p], qh);
        zi[c][p] = fma2(zi[c][p], ql, s);
        zr[c][p] = fma2(zr[c][p], ql, t ^ 0x8000000080000000ull);
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) { u64 a = add2(zr[c][p], one); a = mul2(a, a); m[p] = fma2(zi[c][p], zi[c][p], a); }
      float lg = 0;
#pragma unroll
      for (int g = 0; g < NP; g += 4) {
        u64 pr = m[g];
#pragma unroll
        for (int p = g + 1; p < g + 4 && p < NP; ++p) pr = mul2(pr, m[p]);
        lg += __log2f(lo(pr) * hi(pr));
      }
      acc += lg;
    }
    bit ^= (it >> 3) & 1;
  }
  float r = acc;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int p = 0; p < NP; ++p) r += lo(zr[c][p]) + lo(zi[c][p]);
  if (r == 123.456f) out[0] = r;
}


// dense part with ND units per lane moved to the FP64 pipe: C chains x (NP pairs in packed FP32 + ND units in double)
template <int C, int NP, int ND>
__global__ void k_dense64(float *out, int iters) {
  extern __shared__ float4 rows[];
  double2 *drows = reinterpret_cast<double2 *>(rows + 2 * NP * 32);
  for (int i = threadIdx.x; i < 2 * NP * 32; i += blockDim.x) rows[i] = make_float4(1.0001f, 0.9999f, 1e-4f, -1e-4f);
  for (int i = threadIdx.x; i < 2 * ND * 32; i += blockDim.x) drows[i] = make_double2(1.0001, 1e-4);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  u64 zr[C][NP], zi[C][NP];
  double dr[C][ND], di[C][ND];
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int p = 0; p < NP; ++p) { zr[c][p] = pk(0.1f + lane * 1e-3f, 0.2f); zi[c][p] = pk(0.05f, 0.01f * p); }
#pragma unroll
    for (int d = 0; d < ND; ++d) { dr[c][d] = 0.1 + lane * 1e-3; di[c][d] = 0.05 + 0.01 * d; }
  }
  const u64 one = pk(1.f, 1.f);
  float acc = 0;
  unsigned bit = lane & 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 *rp = rows + ((bit ^ c) & 1) * NP * 32 + lane;
      const double2 *dp = drows + ((bit ^ c) & 1) * ND * 32 + lane;
      u64 m[NP];
      double md = 1.0;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double2 q = dp[d * 32];
        const double s = dr[c][d] * q.y, t = di[c][d] * q.y;
        di[c][d] = fma(di[c][d], q.x, s);
        dr[c][d] = fma(dr[c][d], q.x, -t);
        double a = dr[c][d] + 1.0;
        a = a * a;
        a = fma(di[c][d], di[c][d], a);
        md = d == 0 ? a : md * a;
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const float4 q = rp[p * 32];
        const u64 ql = pk(q.x, q.y), qh = pk(q.z, q.w);
        const u64 s = mul2(zr[c][p], qh), t = mul2(zi[c][p], qh);
        zi[c][p] = fma2(zi[c][p], ql, s);
        zr[c][p] = fma2(zr[c][p], ql, t ^ 0x8000000080000000ull);
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) { u64 a = add2(zr[c][p], one); a = mul2(a, a); m[p] = fma2(zi[c][p], zi[c][p], a); }
      float lg = __log2f((float)md);
#pragma unroll
      for (int g = 0; g < NP; g += 4) {
        u64 pr = m[g];
#pragma unroll
        for (int p = g + 1; p < g + 4 && p < NP; ++p) pr = mul2(pr, m[p]);
        lg += __log2f(lo(pr) * hi(pr));
      }
      acc += lg;
    }
    bit ^= (it >> 3) & 1;
  }
  float r = acc;
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int p = 0; p < NP; ++p) r += lo(zr[c][p]) + lo(zi[c][p]);
#pragma unroll
    for (int d = 0; d < ND; ++d) r += (float)(dr[c][d] + di[c][d]);
  }
  if (r == 123.456f) out[0] = r;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int SM = pr.multiProcessorCount; const double ghz = clk * 1e-6;
  printf("%s, %d SMs, %.3f GHz nominal\n", pr.name, SM, ghz);
  float *out; cudaMalloc(&out, 4);
  const int iters = 20000;
#define RUN_FMA(ILP, PACKED) for (int w : {4, 8, 16, 32}) { \
    float ms = timeit([&] { k_fma<ILP, PACKED><<<SM, w * 32, 0>>>(out, iters, 0.999f); }); \
    double inst = (double)iters * 8 * ILP;  /* per warp */ \
    double cyc = ms * 1e-3 * ghz * 1e9; \
    printf("%s ILP %d warps/SM %2d: %.2f cycles per warp-instruction, %.1f FMA/clk/SM\n", PACKED ? "FFMA2" : "FFMA ", ILP, w, cyc / inst, \
           inst * w * 32 * (PACKED ? 2 : 1) / cyc); }
  RUN_FMA(1, true) RUN_FMA(2, true) RUN_FMA(4, true) RUN_FMA(8, true)
  RUN_FMA(1, false) RUN_FMA(4, false) RUN_FMA(8, false)
#define RUN_DENSE(C, NP, MODE) for (int w : {4, 8, 12, 16}) { \
    cudaFuncSetAttribute(k_dense<C, NP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * NP * 32 * 16); \
    float ms = timeit([&] { k_dense<C, NP, MODE><<<SM, w * 32, 2 * NP * 32 * 16>>>(out, 4000); }); \
    double cyc = ms * 1e-3 * ghz * 1e9; \
    printf("dense mode %d (0 full, 1 no LDS, 2 LDS only) C=%d NP=%d warps/SM %2d: %.0f cycles per warp-step, %.0f cycles per step-round/SMSP, %.1f chain-steps/kclk/SM\n", MODE, C, NP, w, cyc / 4000, cyc / 4000, \
           4000.0 * w * C / cyc * 1000); }
  RUN_DENSE(1, 7, 0) RUN_DENSE(1, 7, 1) RUN_DENSE(1, 7, 2) RUN_DENSE(2, 7, 0) RUN_DENSE(2, 7, 1) RUN_DENSE(2, 7, 2)
#define RUN_D64(C, NP, ND) for (int w : {4, 8, 12, 16}) { \
    const int sm = 2 * NP * 32 * 16 + 2 * ND * 32 * 16; \
    cudaFuncSetAttribute(k_dense64<C, NP, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); \
    float ms = timeit([&] { k_dense64<C, NP, ND><<<SM, w * 32, sm>>>(out, 4000); }); \
    double cyc = ms * 1e-3 * ghz * 1e9; \
    printf("dense64 C=%d: %d pairs FP32 + %d units FP64 per lane, warps/SM %2d: %.0f cycles per warp-step, %.1f chain-steps/kclk/SM\n", C, NP, ND, w, cyc / 4000, \
           4000.0 * w * C / cyc * 1000); }
  RUN_D64(1, 5, 3) RUN_D64(2, 5, 3) RUN_D64(2, 4, 5) RUN_D64(2, 5, 2) RUN_D64(2, 6, 1)
  return 0;
}
