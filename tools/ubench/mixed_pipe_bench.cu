// Do the FP64 and the FP32 pipes of an sm_100 SM run side by side?  One warp instruction stream with NF independent
// fma.rn.f32x2 chains and ND independent fma.rn.f64 chains per iteration; reports cycles per iteration per SM
// sub-partition against the FP32-only and FP64-only streams.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mixed_pipe_bench mixed_pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }

template <int NF, int ND>
__global__ void k_mix(float *out, int iters, float s, double sd) {
  u64 acc[NF > 0 ? NF : 1]; double da[ND > 0 ? ND : 1];
  const u64 b = pk(s, s * 1.0001f), c = pk(1e-9f, 2e-9f);
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = pk(threadIdx.x * 1e-3f + i, i * 0.5f);
#pragma unroll
  for (int i = 0; i < ND; ++i) da[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < (NF > ND ? NF : ND); ++i) {
        if (i < NF) acc[i] = fma2(acc[i], b, c);
        if (i < ND) da[i] = fma(da[i], sd, 1e-9);
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) r += (double)acc[i];
#pragma unroll
  for (int i = 0; i < ND; ++i) r += da[i];
  if (r == 123.456) out[0] = (float)r;
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
  float *out; cudaMalloc(&out, 4);
  const int iters = 20000;
#define RUN(NF, ND) for (int w : {4, 8, 16}) { \
    float ms = timeit([&] { k_mix<NF, ND><<<SM, w * 32, 0>>>(out, iters, 0.999f, 0.999); }); \
    double cyc = ms * 1e-3 * ghz * 1e9 / ((double)iters * 4) / (w / 4.0); \
    printf("FFMA2 x%d + DFMA x%d per round, warps/SM %2d: %.2f cycles per round per warp on its sub-partition (FP32-only would be %d, FP64-only ?)\n", NF, ND, w, cyc, 2 * NF); }
  RUN(8, 0) RUN(0, 8) RUN(0, 4) RUN(8, 4) RUN(8, 8) RUN(8, 2) RUN(6, 3)
  return 0;
}
