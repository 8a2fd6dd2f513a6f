// TMA form of the lean CTA-synchronous sampler (sampler_lean.cuh): identical Metropolis step, different ROW STAGING.
//   * the q rows of a chunk are requested with cp.async.bulk (the TMA unit; SASS UBLKCP): one instruction per 3.5 KB row,
//     no registers, completion counted in bytes on an mbarrier -- instead of 7 LDGSTS per lane and row;
//   * three row buffers with a full / empty mbarrier pair each instead of two buffers and one __syncthreads per chunk: a
//     warp that finishes chunk j releases its buffer, waits until every warp has released chunk j - 1 (a whole chunk
//     ago) and requests chunk j + 2 into that buffer.  No CTA-wide barrier in the steady state: warps may drift a chunk
//     apart instead of meeting every 8 steps (the CTA barrier was 5 % of the stall samples of the cp.async form).
// Kept as a separate kernel: at 128 registers the step body is sensitive to any change of the surrounding code.
#pragma once

#include "sampler_lean.cuh"

namespace nkfb {

// ---- TMA bulk copies + mbarriers (PIPE = 1 form of the row staging)
__device__ __forceinline__ uint32_t lean_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lean_mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void lean_mbar_arrive(uint32_t mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void lean_mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void lean_mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LEAN_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LEAN_DONE;\n\t"
      "bra LEAN_WAIT;\n\t"
      "LEAN_DONE:\n\t"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy by the TMA unit (no registers, one instruction per row), completion counted in bytes on `mbar`
__device__ __forceinline__ void lean_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

inline size_t lean_tma_smem_bytes(int NP, int C, int W, int W32, int nbuf = 3) {
  return (size_t)nbuf * 2 * LEAN_T * 4 * NP * 32 * 4 + (size_t)nbuf * LEAN_T * 16 + (size_t)W * C * LEAN_SB * 4 +
         (size_t)W * C * W32 * 4;
}

// NP pairs of units per lane; C chains per warp; GL pairs per logarithm; UNR unroll factor of the steps of a chunk;
// W_ warps per CTA; SG pairs whose row entries are in registers at a time
// PIPE 0: rows staged with cp.async, two buffers, one __syncthreads per chunk.  PIPE 1: rows staged by the TMA unit
// (cp.async.bulk, one instruction per row), three buffers, full / empty mbarriers per buffer: a warp that finishes
// chunk j releases its buffer, waits until every warp has released chunk j - 1 (a whole chunk ago) and requests
// chunk j + 2 into that buffer -- no CTA-wide barrier in the steady state, warps may drift a chunk apart.
template <int NP, int C, int GL, int UNR, int W_, int SG, int PIPE>
__global__ void __launch_bounds__(32 * W_, 1) sample_ratio_lean_tma_kernel(NetDev<float> net, const float *ws, const float s0,
                                                                   const float s1, SampArgs a, const int *site_seq,
                                                                   const float *a0_seq) {
  using P = P2<float>;
  using V = float2;
  constexpr int L = 32, T = LEAN_T, SB = LEAN_SB;
  constexpr int NG = (NP + GL - 1) / GL;
  constexpr int ROW = 4 * NP * L;  // floats per (table, site) row
  extern __shared__ __align__(16) unsigned char lean_tma_smem[];
  if (!samp_window_ok(a)) return;  // uniform over the grid: taken before any barrier
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  constexpr int NBUF = PIPE ? 3 : 2;
  static_assert(!PIPE || W_ == 2 * LEAN_T, "TMA staging: one row of a chunk per warp");
  float *rows = reinterpret_cast<float *>(lean_tma_smem);                                 // [NBUF][T][2][ROW]
  int4 *steptab = reinterpret_cast<int4 *>(rows + (size_t)NBUF * 2 * T * ROW);       // [NBUF][T]
  int *s_tu = reinterpret_cast<int *>(steptab + NBUF * T) + (size_t)warp * C * SB;   // [C][SB] of this warp
  volatile uint32_t *sigx =
      reinterpret_cast<uint32_t *>(reinterpret_cast<int *>(steptab + NBUF * T) + (size_t)W * C * SB) + warp * C * W32;
  __shared__ __align__(8) uint64_t lean_bars[6];                                       // full[3], empty[3]
  const uint32_t bar_full = lean_smem_u32(lean_bars), bar_empty = lean_smem_u32(lean_bars + 3);
  if (PIPE && threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      lean_mbar_init(bar_full + 8 * b, 2 * LEAN_T);   // one arrive.expect_tx per row of the chunk
      lean_mbar_init(bar_empty + 8 * b, W_);          // one arrive per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (PIPE) __syncthreads();
  const int64_t chain0 = ((int64_t)blockIdx.x * W + warp) * C;
  const V one = P::mk(1.0f, 1.0f);
  const int n_valid = a.n_chains - chain0 >= C ? C : (a.n_chains > chain0 ? (int)(a.n_chains - chain0) : 0);

  // ---- configurations
  for (int idx = lane; idx < C * W32; idx += 32) {
    const int c = idx / W32, w = idx - c * W32;
    uint32_t v = 0;
    if (c < n_valid) {
      if (a.init_random) {
        u32x4 ctr = {(uint32_t)(w >> 2), 0u, (uint32_t)(a.chain_offset + chain0 + c), (uint32_t)TAG_INIT};
        v = sel4(philox4x32_10(ctr, a.k0, a.k1), w & 3);
      } else {
        v = a.chain_state[(chain0 + c) * W32 + w];
      }
      const int nb = N - 32 * w;
      if (nb < 32) v &= (1u << nb) - 1u;
    }
    sigx[idx] = v;
  }
  __syncwarp();

  const int n_sweeps = a.n_discard + a.chain_length;
  const uint32_t n_steps = (uint32_t)n_sweeps * (uint32_t)a.sweep_size;
  const uint32_t phase = (uint32_t)(a.step_offset & (uint64_t)(SB - 1));
  const uint64_t batch0 = a.step_offset - phase;

  // floor(2^20 log2 u) of the batch that starts `pos0` steps after batch0, for the C chains of this warp; a chain slot
  // without a chain gets a threshold nothing reaches
  auto refill = [&](uint32_t pos0) {
    __syncwarp();
    const uint64_t blk = ((batch0 + (uint64_t)pos0) >> 2) + (uint64_t)lane;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      u32x4 ctr = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)(a.chain_offset + chain0 + c), (uint32_t)TAG_UNIFORM};
      const u32x4 ru = philox4x32_10(ctr, a.k0, a.k1);
      int4 t;
      t.x = __float2int_rd(SMath<float>::log2u(ru.x) * LEAN_FT);
      t.y = __float2int_rd(SMath<float>::log2u(ru.y) * LEAN_FT);
      t.z = __float2int_rd(SMath<float>::log2u(ru.z) * LEAN_FT);
      t.w = __float2int_rd(SMath<float>::log2u(ru.w) * LEAN_FT);
      if (c >= n_valid) t = make_int4(1 << 30, 1 << 30, 1 << 30, 1 << 30);
      *reinterpret_cast<int4 *>(s_tu + c * SB + 4 * lane) = t;
    }
    __syncwarp();
  };
  // request the 2 T rows of chunk qn (positions qn T .. qn T + T - 1) into buffer qn & 1, and write its step table
  auto issue_rows = [&](uint32_t qn) {
    for (int r = warp; r < 2 * T; r += W) {
      const int tt = r >> 1, tab = r & 1;
      const int64_t sidx = ((int64_t)qn << LEAN_TSH) + tt - (int64_t)phase;
      const uint32_t sc = sidx < 0 ? 0u : (sidx >= (int64_t)n_steps ? n_steps - 1 : (uint32_t)sidx);
      const int st = __ldg(site_seq + sc);
      const char *src = reinterpret_cast<const char *>(ws + (size_t)(tab * N + st) * ROW);
      char *dst = reinterpret_cast<char *>(rows + ((((qn & 1u) << LEAN_TSH) + tt) * 2 + tab) * ROW);
#pragma unroll
      for (int k = 0; k < NP; ++k) cp_async16(dst + lane * 16 + k * 512, src + lane * 16 + k * 512);  // ROW = 512 NP bytes
    }
    if (threadIdx.x < T) {
      const int tt = threadIdx.x;
      const int64_t sidx = ((int64_t)qn << LEAN_TSH) + tt - (int64_t)phase;
      const uint32_t sc = sidx < 0 ? 0u : (sidx >= (int64_t)n_steps ? n_steps - 1 : (uint32_t)sidx);
      const int st = __ldg(site_seq + sc);
      steptab[((qn & 1u) << LEAN_TSH) + tt] =
          make_int4((st >> 5) * 4, (int)(1u << (st & 31)), __float2int_rn(__ldg(a0_seq + sc) * LEAN_FT), st);
    }
  };

  // PIPE 1: chunk ordinal j (absolute chunk q0 + j) into buffer j % 3: warp w requests row w (step w / 2, table w % 2)
  // with one bulk copy; warp 0 also writes the step table before its arrive (release) on the buffer's full barrier
  auto issue_rows_tma = [&](uint32_t j, uint32_t q0_) {
    const uint32_t b = j % 3u;
    const int tt = warp >> 1, tab = warp & 1;
    const int64_t sidx = ((int64_t)(q0_ + j) << LEAN_TSH) + tt - (int64_t)phase;
    const uint32_t sc = sidx < 0 ? 0u : (sidx >= (int64_t)n_steps ? n_steps - 1 : (uint32_t)sidx);
    const int st = __ldg(site_seq + sc);
    if (warp == 0 && lane < T) {
      const int64_t sl = ((int64_t)(q0_ + j) << LEAN_TSH) + lane - (int64_t)phase;
      const uint32_t scl = sl < 0 ? 0u : (sl >= (int64_t)n_steps ? n_steps - 1 : (uint32_t)sl);
      const int stl = __ldg(site_seq + scl);
      steptab[b * T + lane] =
          make_int4((stl >> 5) * 4, (int)(1u << (stl & 31)), __float2int_rn(__ldg(a0_seq + scl) * LEAN_FT), stl);
    }
    __syncwarp();
    if (lane == 0) {
      const uint32_t dst = lean_smem_u32(rows + ((size_t)(b * T + tt) * 2 + tab) * ROW);
      lean_mbar_arrive_expect_tx(bar_full + 8 * b, (uint32_t)(ROW * sizeof(float)));
      lean_bulk_g2s(dst, ws + (size_t)(tab * N + st) * ROW, (uint32_t)(ROW * sizeof(float)), bar_full + 8 * b);
    }
  };

  V Zr[C][NP], Zi[C][NP];
  float nlg[C];  // 32 - 2^26 (this lane's sum_j log2 |1 + z_j|^2): the constant term of the fixed-point addend
  unsigned n_acc = 0;

  // z <- z * row (in place); returns this lane's sum_j log2 |1 + z_j|^2.  rp: shared memory, lane offset applied.
  auto multiply = [&](V (&zr)[NP], V (&zi)[NP], const float *rp, bool want_log) -> float {
    // SG pairs at a time (row loads, complex multiply, |1 + z|^2), one logarithm per GL pairs
    float acc = 0.0f;
#pragma unroll
    for (int gq = 0; gq < NG; ++gq) {
      V prod = one;
#pragma unroll
      for (int sq = 0; sq < GL; sq += SG) {
        V qlo[SG], qhi[SG], m[SG];
#pragma unroll
        for (int i = 0; i < SG; ++i) {
          const int p = gq * GL + sq + i;
          if (sq + i < GL && p < NP) {
            const float4 q = *reinterpret_cast<const float4 *>(rp + 4 * L * p);
            qlo[i] = make_float2(q.x, q.y);
            qhi[i] = make_float2(q.z, q.w);
          }
        }
#pragma unroll
        for (int i = 0; i < SG; ++i) {
          const int p = gq * GL + sq + i;
          if (sq + i < GL && p < NP) lean_cmul(zr[p], zi[p], qlo[i], qhi[i]);
        }
        if (want_log) {
#pragma unroll
          for (int i = 0; i < SG; ++i) {
            const int p = gq * GL + sq + i;
            if (sq + i < GL && p < NP) m[i] = add2_(zr[p], one);
          }
#pragma unroll
          for (int i = 0; i < SG; ++i) {
            const int p = gq * GL + sq + i;
            if (sq + i < GL && p < NP) m[i] = P::mul(m[i], m[i]);
          }
#pragma unroll
          for (int i = 0; i < SG; ++i) {
            const int p = gq * GL + sq + i;
            if (sq + i < GL && p < NP) m[i] = P::fma(zi[p], zi[p], m[i]);
          }
#pragma unroll
          for (int i = 0; i < SG; ++i) {
            const int p = gq * GL + sq + i;
            if (sq + i < GL && p < NP) prod = (sq + i == 0) ? m[i] : P::mul(prod, m[i]);
          }
        }
      }
      if (want_log) {
        const float lv = P::lg2(prod.x * prod.y);
        acc = gq == 0 ? lv : acc + lv;
      }
    }
    return acc;
  };

  // theta from scratch (pairs of units), then z = exp(-2 theta); padded units carry z = 0
  auto refresh = [&]() {
    const V *cW = reinterpret_cast<const V *>(net.W);
    const V *cb = reinterpret_cast<const V *>(net.b);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int j0 = 2 * L * p + lane, j1 = j0 + L;
      const V b0 = (cb && j0 < M) ? cb[j0] : P::mk(0.0f, 0.0f);
      const V b1 = (cb && j1 < M) ? cb[j1] : P::mk(0.0f, 0.0f);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        Zr[c][p] = P::mk(b0.x, b1.x);
        Zi[c][p] = P::mk(b0.y, b1.y);
      }
    }
    for (int i = 0; i < N; ++i) {
      V xx[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float xv = ((sigx[c * W32 + (i >> 5)] >> (i & 31)) & 1u) ? s1 : s0;
        xx[c] = P::mk(xv, xv);
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int j0 = 2 * L * p + lane, j1 = j0 + L;
        const V w0 = (j0 < M) ? __ldg(cW + (int64_t)i * M + j0) : P::mk(0.0f, 0.0f);
        const V w1 = (j1 < M) ? __ldg(cW + (int64_t)i * M + j1) : P::mk(0.0f, 0.0f);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          Zr[c][p] = P::fma(xx[c], P::mk(w0.x, w1.x), Zr[c][p]);
          Zi[c][p] = P::fma(xx[c], P::mk(w0.y, w1.y), Zi[c][p]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int j0 = 2 * L * p + lane, j1 = j0 + L;
        float e0 = exp_(-2.0f * Zr[c][p].x), e1 = exp_(-2.0f * Zr[c][p].y);
        float sn0, cs0, sn1, cs1;
        sincos_(-2.0f * Zi[c][p].x, &sn0, &cs0);
        sincos_(-2.0f * Zi[c][p].y, &sn1, &cs1);
        if (j0 >= M) e0 = 0.0f;
        if (j1 >= M) e1 = 0.0f;
        Zr[c][p] = P::mk(e0 * cs0, e1 * cs1);
        Zi[c][p] = P::mk(e0 * sn0, e1 * sn1);
      }
      float acc = 0.0f;
#pragma unroll
      for (int gq = 0; gq < NG; ++gq) {
        V prod = one;
#pragma unroll
        for (int p = gq * GL; p < (gq + 1) * GL && p < NP; ++p) {
          const V aa = add2_(Zr[c][p], one);
          V m = P::mul(aa, aa);
          m = P::fma(Zi[c][p], Zi[c][p], m);
          prod = P::mul(prod, m);
        }
        acc += P::lg2(prod.x * prod.y);
      }
      nlg[c] = fmaf(acc, -LEAN_FX, 32.0f);
    }
  };

  // one Metropolis step of the C chains: table entry `e` (this step), rows at `slot` (lane offset applied), log2 u at
  // s_tu[.][tb]
  auto step = [&](const int4 e, const float *slot, const int tb) {
    const uint32_t msk = (uint32_t)e.y;
    uint32_t word[C];
    bool up[C];
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c)
      word[c] = *reinterpret_cast<volatile uint32_t *>(reinterpret_cast<volatile char *>(sigx + c * W32) + e.x);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      up[c] = (word[c] & msk) != 0u;
      acc[c] = multiply(Zr[c], Zi[c], slot + (up[c] ? ROW : 0), true);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      // accept iff sum over lanes of 2^20 (log2|1+z'|^2 - log2|1+z|^2) > 2^20 (log2 u - (+-A_site))
      int v = __float2int_rn(fmaf(acc[c], LEAN_FX, nlg[c])) >> 6;
      const int thr = s_tu[c * SB + tb] - (up[c] ? -e.z : e.z);
      if (lane == 0) v -= thr;
      if (__builtin_expect(__reduce_add_sync(0xffffffffu, v) > 0, 1)) {
        nlg[c] = fmaf(acc[c], -LEAN_FX, 32.0f);
        ++n_acc;
        *reinterpret_cast<volatile uint32_t *>(reinterpret_cast<volatile char *>(sigx + c * W32) + e.x) = word[c] ^ msk;
      } else {
        // rejected: multiply back by the other row of the slot (the inverse of the one just applied)
        multiply(Zr[c], Zi[c], slot + (up[c] ? 0 : ROW), false);
      }
    }
  };

  // ---- prologue: rows of the first two chunks, random numbers of the first batch
  const uint32_t q0 = phase >> LEAN_TSH;
  const uint32_t n_chunks = ((phase + n_steps + T - 1) >> LEAN_TSH) - q0;   // chunks that hold steps of this call
  if constexpr (PIPE) {
    issue_rows_tma(0, q0);
    if (n_chunks > 1) issue_rows_tma(1, q0);
  } else {
    issue_rows(q0);
    cp_async_wait_all();
    __syncthreads();
    issue_rows(q0 + 1);
  }
  refill(0);
  const float *lane_rows = rows + 4 * lane;
  bool need_refresh = true;             // z is (re)computed from the configuration at the top of a chunk
  uint32_t s = 0;                       // steps done
  int left = a.sweep_size;              // steps left in the current sweep
  int sweep = 0;
  for (uint32_t q = q0; s < n_steps; ++q) {
    const uint32_t pos_q = q << LEAN_TSH;                    // position of the chunk's first slot
    if ((pos_q & (SB - 1)) == 0 && q != q0) refill(pos_q);
    if (need_refresh) {
      refresh();
      need_refresh = false;
    }
    const uint32_t jq = q - q0;                            // chunk ordinal
    const uint32_t buf = PIPE ? jq % 3u : (q & 1u);
    if constexpr (PIPE) lean_mbar_wait(bar_full + 8 * buf, (jq / 3u) & 1u);   // rows and step table of this chunk
    const float *cbase = lane_rows + (size_t)(buf << LEAN_TSH) * (2 * ROW);
    const int4 *tbase = steptab + (buf << LEAN_TSH);
    const int tb0 = (int)(pos_q & (SB - 1));
    const bool whole = pos_q >= phase && (uint64_t)pos_q - phase + T <= n_steps && left >= T;
    if (whole) {
#pragma unroll UNR
      for (int tt = 0; tt < T; ++tt) step(tbase[tt], cbase + tt * (2 * ROW), tb0 + tt);
      s += T;
      left -= T;
    } else {
      for (int tt = 0; tt < T; ++tt) {
        const int64_t sidx = (int64_t)pos_q + tt - (int64_t)phase;
        if (sidx < 0 || sidx >= (int64_t)n_steps) continue;
        step(tbase[tt], cbase + tt * (2 * ROW), tb0 + tt);
        ++s;
        if (--left == 0 && tt != T - 1) {
          // sweep boundary inside the chunk: emit the configurations
          left = a.sweep_size;
          if (sweep >= a.n_discard && lane < W32) {
#pragma unroll
            for (int c = 0; c < C; ++c)
              if (c < n_valid)
                a.samples[((chain0 + c) * a.chain_length + (sweep - a.n_discard)) * W32 + lane] = sigx[c * W32 + lane];
          }
          ++sweep;
          if ((sweep % RATIO_REFRESH) == 0) need_refresh = true;
        }
      }
    }
    if (left == 0) {
      left = a.sweep_size;
      if (sweep >= a.n_discard && lane < W32) {
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (c < n_valid)
            a.samples[((chain0 + c) * a.chain_length + (sweep - a.n_discard)) * W32 + lane] = sigx[c * W32 + lane];
      }
      ++sweep;
      if ((sweep % RATIO_REFRESH) == 0) need_refresh = true;
    }
    if constexpr (PIPE) {
      // release this chunk's buffer; once every warp has released the PREVIOUS chunk's, request chunk j + 2 into it
      __syncwarp();
      if (lane == 0) lean_mbar_arrive(bar_empty + 8 * buf);
      if (jq + 2 < n_chunks) {
        if (jq >= 1) lean_mbar_wait(bar_empty + 8 * ((jq - 1) % 3u), ((jq - 1) / 3u) & 1u);
        issue_rows_tma(jq + 2, q0);
      }
    } else {
      // end of a chunk: the next one has landed everywhere, this one is free for the one after it
      cp_async_wait_all();
      __syncthreads();
      issue_rows(q + 2);
    }
  }
  if (!PIPE) cp_async_wait_all();
  if (lane < W32) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (c < n_valid) a.chain_state[(chain0 + c) * W32 + lane] = sigx[c * W32 + lane];
  }
  if (a.n_accepted && lane == 0 && n_acc) atomicAdd(a.n_accepted, (unsigned long long)n_acc);
}

template <int NP, int C, int GL, int UNR = 8, int W_ = 16, int SG = 2, int PIPE = 1>
static int launch_ratio_lean_tma(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, const void *ws,
                             const int *site_seq, const float *a0_seq, cudaStream_t st) {
  constexpr int W = W_;
  const size_t smem = lean_tma_smem_bytes(NP, C, W, (net->n_visible + 31) >> 5, 3);
  cudaError_t rc = cudaFuncSetAttribute(sample_ratio_lean_tma_kernel<NP, C, GL, UNR, W_, SG, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem);
  if (rc != cudaSuccess) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "sampler needs %zu B of shared memory", smem);
  const int64_t blocks = (a.n_chains + (int64_t)C * W - 1) / ((int64_t)C * W);
  sample_ratio_lean_tma_kernel<NP, C, GL, UNR, W_, SG, PIPE><<<(unsigned)blocks, W * 32, smem, st>>>(
      make_netdev<float>(net), (const float *)ws, (float)a.s0, (float)a.s1, a, site_seq, a0_seq);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
