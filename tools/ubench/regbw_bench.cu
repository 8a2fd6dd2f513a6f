// Microbenchmarks behind the round-2 sampler work (sm_100a):
//  A. fma.rn.f32x2 issue rate as a function of how many DISTINCT 64-bit register operands an instruction reads
//     (acc = acc * b + c with b, c shared by all accumulators  vs  every accumulator its own b_i, c_i);
//  B. shared-memory loads: LDS.128 with per-lane addresses (512 B per warp), LDS.128 broadcast (all lanes one
//     address), LDS.64, LDS.32 -- cycles per load instruction per SM with 16 warps;
//  C. FFMA2 next to SEL (ALU pipe): 8 FFMA2 + K SEL per round.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o regbw_bench regbw_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pk(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }

// MODE 0: a_i = a_i * b + c (b, c shared); 1: a_i = a_i * b_i + c_i; 2: a_i = a_i * b_i (mul, two distinct);
// 3: a_i = b_i * c_i + a_i (addend is the accumulator); 4: t = a_i * b_i ; a_i = a_i * c_i + t  (the cmul pattern)
template <int ILP, int MODE>
__global__ void k_ops(float *out, int iters, float s) {
  u64 a[ILP], b[ILP], c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    a[i] = pk(threadIdx.x * 1e-3f + i, i * 0.5f);
    b[i] = pk(s + 1e-6f * i, s * 1.0001f);
    c[i] = pk(1e-9f * (i + 1), 2e-9f);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) a[i] = fma2(a[i], b[0], c[0]);
        if (MODE == 1) a[i] = fma2(a[i], b[i], c[i]);
        if (MODE == 2) a[i] = mul2(a[i], b[i]);
        if (MODE == 3) a[i] = fma2(b[i], c[i], a[i]);
        if (MODE == 4) { const u64 t = mul2(a[i], b[i]); a[i] = fma2(a[i], c[i], t); }
      }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += (float)(a[i] & 0xffff) + (float)(b[i] & 1) + (float)(c[i] & 1);
  if (r == 123.456f) out[0] = r;
}

// MODE 0: LDS.128 per-lane; 1: LDS.128 broadcast; 2: LDS.64 per-lane; 3: LDS.32 per-lane; 4: LDS.32 broadcast;
// 5: LDS.64 broadcast
template <int MODE>
__global__ void k_lds(float *out, int iters) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float acc = 0;
  unsigned off = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const unsigned base = (off + r * 32) & 1023;
      if (MODE == 0) { const float4 v = sm[base + lane]; acc += (v.x + v.y) + (v.z + v.w); }
      if (MODE == 1) { const float4 v = sm[base]; acc += (v.x + v.y) + (v.z + v.w); }
      if (MODE == 2) { const float2 v = reinterpret_cast<const float2 *>(sm)[base + lane]; acc += v.x + v.y; }
      if (MODE == 3) { const float v = reinterpret_cast<const float *>(sm)[base + lane]; acc += v; }
      if (MODE == 4) { const float v = reinterpret_cast<const float *>(sm)[base]; acc += v; }
      if (MODE == 5) { const float2 v = reinterpret_cast<const float2 *>(sm)[base]; acc += v.x + v.y; }
    }
    off += 256 + (acc == 1e30f);
  }
  if (acc == 123.456f) out[0] = acc;
}

// 8 FFMA2 (distinct multiplicands) + K SEL on a warp-uniform predicate per round
template <int K>
__global__ void k_sel(float *out, int iters, float s, int flag) {
  u64 a[8], b[8];
  float qa[8], qb[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = pk(threadIdx.x * 1e-3f + i, i * 0.5f);
    b[i] = pk(s + 1e-6f * i, s);
    qa[i] = s + i;
    qb[i] = s - i;
    q[i] = 0;
  }
  for (int it = 0; it < iters; ++it) {
    const bool p = ((it ^ flag) & 1) != 0;
#pragma unroll
    for (int i = 0; i < K; ++i) q[i] = p ? qa[i] : qb[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], b[i], pk(q[i % (K > 0 ? K : 1)], q[(i + 1) % (K > 0 ? K : 1)]));
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += (float)(a[i] & 0xffff);
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
#define RUN_OPS(ILP, MODE, NI) for (int w : {4, 8, 16}) { \
    float ms = timeit([&] { k_ops<ILP, MODE><<<SM, w * 32, 0>>>(out, iters, 0.999f); }); \
    double inst = (double)iters * 4 * ILP * NI; double cyc = ms * 1e-3 * ghz * 1e9; \
    printf("ops mode %d ILP %d warps/SM %2d: %.2f cycles per packed instruction per sub-partition\n", MODE, ILP, w, cyc / inst / (w / 4.0)); }
  RUN_OPS(8, 0, 1) RUN_OPS(8, 1, 1) RUN_OPS(8, 2, 1) RUN_OPS(8, 3, 1) RUN_OPS(8, 4, 2)
#define RUN_LDS(MODE) for (int w : {4, 16}) { \
    float ms = timeit([&] { k_lds<MODE><<<SM, w * 32, 32768>>>(out, iters); }); \
    double inst = (double)iters * 8 * w; double cyc = ms * 1e-3 * ghz * 1e9; \
    printf("lds mode %d (0 128b lane, 1 128b bcast, 2 64b lane, 3 32b lane, 4 32b bcast, 5 64b bcast) warps/SM %2d: %.2f cycles per load instruction per SM\n", MODE, w, cyc / inst); }
  RUN_LDS(0) RUN_LDS(1) RUN_LDS(2) RUN_LDS(3) RUN_LDS(4) RUN_LDS(5)
#define RUN_SEL(K) for (int w : {4, 16}) { \
    float ms = timeit([&] { k_sel<K><<<SM, w * 32, 0>>>(out, iters, 0.999f, 0); }); \
    double cyc = ms * 1e-3 * ghz * 1e9; \
    printf("8 FFMA2 + %d SEL per round, warps/SM %2d: %.2f cycles per round per sub-partition (FFMA2 alone: 16)\n", K, w, cyc / iters / (w / 4.0)); }
  RUN_SEL(0) RUN_SEL(4) RUN_SEL(8)
  return 0;
}
