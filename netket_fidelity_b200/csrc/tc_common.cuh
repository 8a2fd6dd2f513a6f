// Shared pieces of the tensor-core (tcgen05 + TMEM) kernels: geometry, PTX wrappers, operand preparation.
#pragma once
#include <cstdio>
#include <cstring>
#include <algorithm>

#include <cuda_bf16.h>

#include "tile.cuh"

namespace nkfb {

constexpr int TC_THREADS = 256;
constexpr int TC_ROWS = 128;  // sample pairs per tile (UMMA M)
constexpr int TC_T = 2;       // tiles sharing one staged chunk of W~

struct TcGeom {
  int N, M, KT, KC, KS, NC, NN;  // KT = padded K, KC = KT/8 core chunks, KS = KT/16 k-steps, NC chunks of NN columns
  __host__ __device__ size_t b_split_bytes() const { return (size_t)KC * NN * 16; }
  __host__ __device__ size_t b_chunk_bytes() const { return 3 * b_split_bytes(); }
  __host__ __device__ size_t a_bytes() const { return (size_t)KC * TC_ROWS * 16; }
  __host__ __device__ size_t prep_bytes_per_net() const { return (size_t)NC * b_chunk_bytes(); }
};

// cols: real columns to cover; extra_col_bytes: additional shared memory per column (pass B keeps the three
// bf16 splits of a 128-sample Y tile); x_bytes: bytes of the X tile
inline bool tc_geometry(int N, int M, TcGeom *g, int cols, size_t extra_col_bytes = 0, size_t x_bytes = 0,
                        size_t budget_bytes = 200 * 1024, int nn_quantum = 16, int nn_cap = 256) {
  g->N = N;
  g->M = M;
  g->KT = ((N + 1 + 15) / 16) * 16;
  g->KC = g->KT / 8;
  g->KS = g->KT / 16;
  if (x_bytes == 0) x_bytes = (size_t)g->KC * TC_ROWS * 16;
  // columns per chunk: multiple of 16, <= 256, and everything must fit in shared memory
  if (budget_bytes < x_bytes + 8192 + 4096) return false;
  const size_t budget = budget_bytes - x_bytes - 8192;
  int nn_max = (int)std::min<size_t>((size_t)nn_cap, budget / ((size_t)3 * g->KC * 16 + extra_col_bytes));
  nn_max = (nn_max / nn_quantum) * nn_quantum;
  if (nn_max < nn_quantum) return false;
  g->NC = (cols + nn_max - 1) / nn_max;
  g->NN = (((cols + g->NC - 1) / g->NC) + nn_quantum - 1) / nn_quantum * nn_quantum;
  return g->KT <= 1024;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (sm_100): start address [0,14), leading byte offset [16,30), stride byte offset [32,46)
  // (all >> 4), version = 1 at [46,48), base offset 0, layout type SWIZZLE_NONE = 0 at [61,64)
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int Mdim, int Ndim, int a_mn_major, int b_mn_major) {
  // InstrDescriptor: c_format F32 = 1 at [4,6); a_format BF16 = 1 at [7,10); b_format BF16 = 1 at [10,13);
  // a_major [15], b_major [16] (0 = K-major); n_dim = N >> 3 at [17,23); m_dim = M >> 4 at [24,29)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(Ndim >> 3) << 17) | ((uint32_t)(Mdim >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
// NKFB_MBAR_HINT (ns): upper bound of the time a try_wait may stay suspended before it returns false; the default
// interval made the waiting epilogue warps poll (SYNCS + YIELD + BRA were a fifth of the executed instructions)
#ifndef NKFB_MBAR_HINT
#define NKFB_MBAR_HINT 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
#if NKFB_MBAR_HINT > 0
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(mbar),
      "r"(parity), "r"((uint32_t)NKFB_MBAR_HINT)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
#endif
}
// the same wait for a warp that has nothing else to do (the UMMA-issuing warp): sleep between polls so that its spin
// does not take issue slots from the epilogue warps of its sub-partition (a tight try_wait loop ran at 0.22 IPC)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra RDONE;\n\t"
      "RWAIT_LOOP:\n\t"
      "nanosleep.u32 128;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra RDONE;\n\t"
      "bra RWAIT_LOOP;\n\t"
      "RDONE:\n\t"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 8 consecutive fp32 columns of this thread's TMEM lane (row)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// store 8 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 16 consecutive fp32 columns: one wait for two loads
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ operand preparation
// Bprep[chunk][split][kc][n][8] bf16: element (k = 8 kc + e, n) of split `split` of W~, W~[k][col] with
// col = chunk * NN + n: rows k < N from the kernel (N, M) complex interleaved, row N the bias, else 0.
static __global__ void tc_prep_kernel(NetDev<float> net, TcGeom g, __nv_bfloat16 *out) {
  const int64_t total = (int64_t)g.NC * g.KC * g.NN;  // one thread per (chunk, kc, n): 8 k values x 3 splits
  const int cols = 2 * g.M;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % g.NN);
    const int kc = (int)((idx / g.NN) % g.KC);
    const int chunk = (int)(idx / ((int64_t)g.NN * g.KC));
    const int col = chunk * g.NN + n;
    __align__(16) __nv_bfloat16 hi[8], mid[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      float w = 0.f;
      if (col < cols) {
        if (k < g.N)
          w = net.W[(int64_t)k * cols + col];
        else if (k == g.N && net.b)
          w = net.b[col];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      const float r1 = w - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(m);
      hi[e] = h;
      mid[e] = m;
      lo[e] = __float2bfloat16_rn(r2);
    }
    const size_t split_elems = (size_t)g.KC * g.NN * 8;
    __nv_bfloat16 *base = out + (size_t)chunk * 3 * split_elems + ((size_t)kc * g.NN + n) * 8;
    *reinterpret_cast<uint4 *>(base) = *reinterpret_cast<const uint4 *>(hi);
    *reinterpret_cast<uint4 *>(base + split_elems) = *reinterpret_cast<const uint4 *>(mid);
    *reinterpret_cast<uint4 *>(base + 2 * split_elems) = *reinterpret_cast<const uint4 *>(lo);
  }
}

// X tile in the canonical K-major no-swizzle layout [kc][row][8] bf16.
// TC_THREADS = 2 TC_ROWS: thread t always serves row t % 128 and the 8-spin groups kc = t / 128, t / 128 + 2, ...;
// for N <= 128 it reads its row's packed words ONCE (they used to be re-read per group: 7 dependent L2 round
// trips per tile at C4, 13 % of the estimator's time on the long scoreboard) and cuts the groups out of registers.
// (single_load = false keeps the per-group loads: with two CTAs per SM the estimator hides that latency and the
// shorter code wins there -- measured.)
// xtab (optional, shared memory, tc_fill_x_table): the 16 bytes of every group of 8 spins, indexed by its 8 bits --
// one LDS.128 instead of eight shift / test / select rounds for the groups that lie entirely below N (the bit
// arithmetic was a fifth of the instructions of the tensor-core kernels).
__device__ __forceinline__ void tc_fill_x_table(uint4 *xtab, float s0, float s1) {
  const uint32_t b0 = __bfloat16_as_ushort(__float2bfloat16_rn(s0)), b1 = __bfloat16_as_ushort(__float2bfloat16_rn(s1));
  for (int b = threadIdx.x; b < 256; b += blockDim.x) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      w[q] = (((b >> (2 * q)) & 1) ? b1 : b0) | ((((b >> (2 * q + 1)) & 1) ? b1 : b0) << 16);
    xtab[b] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ void tc_build_x_tile(unsigned char *As, const uint32_t *packed, int64_t s_base, int64_t S,
                                                const TcGeom &g, float s0, float s1, bool single_load = false,
                                                const uint4 *xtab = nullptr) {
  const int W32 = (g.N + 31) >> 5;
  const uint16_t b0 = __bfloat16_as_ushort(__float2bfloat16_rn(s0)), b1 = __bfloat16_as_ushort(__float2bfloat16_rn(s1));
  auto emit = [&](int idx, int k0, uint32_t bits, bool live) {
    if (xtab != nullptr && k0 + 8 <= g.N) {
      *reinterpret_cast<uint4 *>(As + (size_t)idx * 16) = live ? xtab[bits] : make_uint4(0, 0, 0, 0);
      return;
    }
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t lo16 = 0, hi16 = 0;
      const int ka = k0 + 2 * q, kb = ka + 1;
      if (live) {
        lo16 = ka < g.N ? (((bits >> (2 * q)) & 1u) ? b1 : b0) : (ka == g.N ? 0x3F80u : 0u);
        hi16 = kb < g.N ? (((bits >> (2 * q + 1)) & 1u) ? b1 : b0) : (kb == g.N ? 0x3F80u : 0u);
      }
      w[q] = lo16 | (hi16 << 16);
    }
    *reinterpret_cast<uint4 *>(As + (size_t)idx * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  };
  if (single_load && W32 <= 4 && g.KC <= 16) {
    const int r = threadIdx.x & (TC_ROWS - 1), kc0 = threadIdx.x / TC_ROWS;
    const int64_t gs = s_base + r;
    const bool live = gs < S;
    uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int w = 0; w < 4; ++w)
      if (live && w < W32) wd[w] = __ldg(packed + gs * W32 + w);
    if (blockDim.x == 4 * TC_ROWS) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {  // 512 threads: kc = kc0 + 4 m, first spin 8 kc0 + 32 m: packed word m
        const int kc = kc0 + 4 * m;
        if (kc < g.KC) {
          const int k0 = kc * 8;
          emit(kc * TC_ROWS + r, k0, (wd[m] >> (k0 & 31)) & 0xFFu, live);
        }
      }
      return;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {  // kc = kc0 + 2 m, first spin 8 kc0 + 16 m: packed word m / 2
      const int kc = kc0 + 2 * m;
      if (kc < g.KC) {
        const int k0 = kc * 8;
        emit(kc * TC_ROWS + r, k0, (wd[m >> 1] >> (k0 & 31)) & 0xFFu, live);
      }
    }
    return;
  }
  for (int idx = threadIdx.x; idx < g.KC * TC_ROWS; idx += blockDim.x) {
    const int kc = idx / TC_ROWS, r = idx - kc * TC_ROWS;
    const int64_t gs = s_base + r;
    uint32_t bits = 0;
    const int k0 = kc * 8;
    if (gs < S && k0 < g.N) bits = (packed[gs * W32 + (k0 >> 5)] >> (k0 & 31)) & 0xFFu;
    emit(idx, k0, bits, gs < S);
  }
}


// Lean form of the builder for N <= 128 (W32 <= 4, KC <= 16) and NQ * 128 threads: thread t serves row t % 128 and the
// 8-spin groups kc = t / 128 + NQ m.  One 16-byte (or W32 4-byte) load of the packed row, then per group a byte
// extraction whose word index and shift are compile-time constants plus 8 (t / 128), one LDS.128 from the spin-group
// table and one STS.128.  The group that holds the bias row (kc == N / 8: the last N % 8 spins, 1.0 at k = N, zeros
// behind) is patched with two 16-byte masks (xtail[0] = AND mask, xtail[1] = OR pattern, tc_fill_x_tail); groups beyond
// it are zero.  Every condition is uniform over a warp.  (The generic loop above spent ~420 instructions per thread
// and tile on index arithmetic: a quarter of the estimator's instructions, ncu source view, profiles/r02_*.)
__device__ __forceinline__ void tc_fill_x_tail(uint4 *xtail, int N) {
  if (threadIdx.x < 8) {
    const int e = threadIdx.x, r = N & 7;  // element e of the bias group: spins for e < r, 1.0 at e == r, zero behind
    uint16_t *am = reinterpret_cast<uint16_t *>(xtail), *om = am + 8;
    am[e] = e < r ? 0xFFFFu : 0u;
    om[e] = e == r ? 0x3F80u : 0u;
  }
}

template <int NQ>
__device__ __forceinline__ void tc_build_x_tile_fast(unsigned char *As, const uint32_t *packed, int64_t s_base, int64_t S,
                                                     const TcGeom &g, const uint4 *xtab, const uint4 *xtail) {
  const int r = threadIdx.x & (TC_ROWS - 1), q = threadIdx.x / TC_ROWS;
  const int W32 = (g.N + 31) >> 5, kt = g.N >> 3;
  const int64_t gs = s_base + r;
  uint4 *dst = reinterpret_cast<uint4 *>(As) + r;
  if (gs >= S) {  // rows beyond the batch (last tile only): zero
#pragma unroll
    for (int m = 0; m < 16 / NQ; ++m)
      if (q + NQ * m < g.KC) dst[(q + NQ * m) * TC_ROWS] = make_uint4(0, 0, 0, 0);
    return;
  }
  uint32_t wd[4] = {0u, 0u, 0u, 0u};
  if (W32 == 4 && (reinterpret_cast<uintptr_t>(packed) & 15u) == 0) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(packed + gs * 4));
    wd[0] = v.x, wd[1] = v.y, wd[2] = v.z, wd[3] = v.w;
  } else {
#pragma unroll
    for (int w = 0; w < 4; ++w)
      if (w < W32) wd[w] = __ldg(packed + gs * W32 + w);
  }
  const uint4 tand = xtail[0], tor = xtail[1];
#pragma unroll
  for (int m = 0; m < 16 / NQ; ++m) {
    const int kc = q + NQ * m;
    if (kc < g.KC) {
      const uint32_t bits = (wd[(NQ * m) >> 2] >> ((((NQ * m) & 3) + q) * 8)) & 0xFFu;
      uint4 e = xtab[bits];
      if (kc >= kt) {
        if (kc == kt)
          e = make_uint4((e.x & tand.x) | tor.x, (e.y & tand.y) | tor.y, (e.z & tand.z) | tor.z, (e.w & tand.w) | tor.w);
        else
          e = make_uint4(0, 0, 0, 0);
      }
      dst[kc * TC_ROWS] = e;
    }
  }
}

inline bool bf16_exact(double v) {
  const float f = (float)v;
  uint32_t u;
  memcpy(&u, &f, 4);
  return (double)f == v && (u & 0xFFFFu) == 0;
}

}  // namespace nkfb
