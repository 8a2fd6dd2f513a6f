// Half-warp form of the lean CTA-synchronous sampler (sampler_lean_tma.cuh): the same Metropolis step, Philox streams,
// site sequence, fixed-point decision and TMA row staging -- but a chain lives on the 16 lanes of a HALF warp, so the
// two chains of a warp advance side by side in ONE instruction stream instead of one after the other.
//
// Why (profiles/r02_sampler_notes.md, r02_sass_hist_*.md): the kernel is bound by the issue port of the SM
// sub-partition (a packed FFMA2 / FMUL2 / FADD2 holds it two cycles, anything else one).  With a chain on 32 lanes, M = 400
// hidden units are 200 pairs on 7 x 32 = 224 slots (10.7 % padding) and every per-step instruction that is not the
// packed arithmetic (table entry, configuration word, bit test, row select, logarithms, fixed-point addend, threshold,
// REDUX, accept bookkeeping) is issued once PER CHAIN: 254 instructions per warp-step of two chains, 108 of them packed.
// On 16 lanes the 200 pairs are 13 x 16 = 208 slots (3.8 % padding: 104 packed instructions per warp-step) and the
// bookkeeping is issued once PER WARP-STEP for both chains: the two halves differ only in data (their configuration
// word, flip direction -> row address, threshold), never in the instruction stream, except when exactly one of them
// rejects its move (the multiply-back by the inverse row then runs on one half).
//   * q tables in the 16-lane layout [dir][site][pair p][lane16]{Re q_j0, Re q_j1, Im q_j0, Im q_j1},
//     j0 = 32 p + lane16, j1 = j0 + 16 (prep_ratio_params with L = 16), ROW = 64 NP floats;
//   * segmented sum of the fixed-point addends: two integer REDUX over the full warp, the other half contributing 0.
#pragma once

#include "sampler_lean_tma.cuh"

namespace nkfb {

// shared-memory accesses by 32-bit shared address (volatile: ordered after the mbarrier waits, never merged; ptxas still
// schedules them among the arithmetic): the generic-pointer form recomputed the shared window base every step
__device__ __forceinline__ float4 hw_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 hw_lds128i(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t hw_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void hw_sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}

// mbarrier wait whose try_wait may stay suspended for up to ~1 ms: the default poll interval made the waiting warps
// issue 7 instructions per Metropolis step (ncu source view), taken from the issue port the other warps need
__device__ __forceinline__ void hw_mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "HW_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra HW_DONE;\n\t"
      "bra HW_WAIT;\n\t"
      "HW_DONE:\n\t"
      "}\n" ::"r"(mbar),
      "r"(parity), "r"(1000000u)
      : "memory");
}

// ints per chain in the per-warp block of shared memory: LEAN_SB thresholds, then the configuration words (one base
// register addresses both)
inline int lean_hw_chain_ints(int W32) { return LEAN_SB + ((W32 + 3) & ~3); }

inline size_t lean_hw_smem_bytes(int NP, int W, int W32) {
  return (size_t)3 * 2 * LEAN_T * 4 * NP * 16 * 4 + (size_t)3 * LEAN_T * 16 + (size_t)W * 2 * lean_hw_chain_ints(W32) * 4;
}

// NP pairs of units per lane (M <= 32 NP); GL pairs per logarithm; UNR unroll factor of the steps of a chunk;
// W_ warps per CTA (= rows of a chunk: one bulk copy per warp); SG pairs whose row entries are in registers at a time
// WA <= W_ warps of the CTA carry chains; the others exit at once (a CTA of fewer chains whose warps are still laid
// out like those of a 16-warp CTA)
// PF = 1 (experiment, off): table entry, configuration word and threshold of step t + 1 are loaded during the multiply
// phase of step t (the word patched if step t flips a bit of the same word), which takes two dependent shared-memory
// round trips off the critical path of a step -- measured SLOWER (5.46 vs 5.11 ms per wave with 16 warps, 3.06 vs 3.00 ms
// with one warp per scheduler, gpurun_out/hw_check7.log): five more live registers cost more than the ~50 cycles saved.
template <int NP, int GL, int UNR, int W_, int SG, int WA = W_, int PF = 0>
__global__ void __launch_bounds__(32 * W_, 1) sample_ratio_lean_hw_kernel(NetDev<float> net, const float *ws, const float s0,
                                                                  const float s1, SampArgs a, const int *site_seq,
                                                                  const float *a0_seq) {
  using P = P2<float>;
  using V = float2;
  constexpr int L = 16, C = 2, T = LEAN_T, SB = LEAN_SB;
  constexpr int NG = (NP + GL - 1) / GL;
  constexpr int ROW = 4 * NP * L;  // floats per (table, site) row
  extern __shared__ __align__(16) unsigned char lean_hw_smem[];
  if (!samp_window_ok(a)) return;  // uniform over the grid: taken before any barrier
  samp_resolve_offset(a);
  const int N = net.N, M = net.M, W32 = (N + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int W = WA;
  const int hl = lane & 15, hc = lane >> 4;  // lane within the half warp; chain slot of this half
  static_assert(W_ >= 8 && W_ <= 32, "warps per CTA");
  float *rows = reinterpret_cast<float *>(lean_hw_smem);                          // [3][T][2][ROW]
  int4 *steptab = reinterpret_cast<int4 *>(rows + (size_t)3 * 2 * T * ROW);       // [3][T]
  const int CB = SB + ((W32 + 3) & ~3);                                          // ints per chain block
  int *s_tu = reinterpret_cast<int *>(steptab + 3 * T) + (size_t)warp * C * CB;   // [C][CB]: thresholds | configuration
  volatile uint32_t *sigx = reinterpret_cast<volatile uint32_t *>(s_tu) + SB;     // chain c: sigx[c * CB + word]
  const uint32_t tu_my = lean_smem_u32(s_tu + hc * CB);                           // shared address of this half's chain block
  volatile uint32_t *sig_my_p = sigx + hc * CB;
  int m_hi = hc ? -1 : 0;  // all ones on the upper half warp; opaque, so that `v & m_hi` stays one LOP3 on a register
  asm volatile("" : "+r"(m_hi));  // instead of a select on a predicate recomputed from the lane id every step
  __shared__ __align__(8) uint64_t hw_bars[6];  // full[3], empty[3]
  const uint32_t bar_full = lean_smem_u32(hw_bars), bar_empty = lean_smem_u32(hw_bars + 3);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      lean_mbar_init(bar_full + 8 * b, 2 * LEAN_T);  // one arrive.expect_tx per row of the chunk
      lean_mbar_init(bar_empty + 8 * b, WA);         // one arrive per warp that carries chains
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (WA < W_ && warp >= WA) return;
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
    sigx[c * CB + w] = v;
  }
  __syncwarp();

  const int n_sweeps = a.n_discard + a.chain_length;
  const uint32_t n_steps = (uint32_t)n_sweeps * (uint32_t)a.sweep_size;
  const uint32_t phase = (uint32_t)(a.step_offset & (uint64_t)(SB - 1));
  const uint64_t batch0 = a.step_offset - phase;

  const uint32_t q0c = phase >> LEAN_TSH;            // first chunk of this call
  const uint32_t rows_addr = lean_smem_u32(rows);

  // floor(2^20 log2 u) of the batch that starts `pos0` steps after batch0, for the two chains of this warp (all 32
  // lanes work for each chain: four uniforms per lane); a chain slot without a chain gets a threshold nothing reaches
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
      *reinterpret_cast<int4 *>(s_tu + c * CB + 4 * lane) = t;
    }
    __syncwarp();
  };

  // chunk ordinal j (absolute chunk q0 + j) into buffer b = j % 3: warp w requests row w (step w / 2, table w % 2) with
  // one bulk copy; warp 0 also writes the step table before its arrive (release) on the buffer's full barrier.
  // Table entry: byte offset of the configuration word inside the chain block, bit mask, A_site in fixed point.
  auto issue_rows_tma = [&](uint32_t j, uint32_t b) {
    const int pos = (int)((q0c + j) << LEAN_TSH) - (int)phase;   // step index of the chunk's first slot
    const int last = (int)n_steps - 1;
    if (warp == 0 && lane < T) {
      const int scl = min(max(pos + lane, 0), last);
      const int stl = __ldg(site_seq + scl);
      steptab[b * T + lane] = make_int4(SB * 4 + (stl >> 5) * 4, (int)(1u << (stl & 31)),
                                        __float2int_rn(__ldg(a0_seq + scl) * LEAN_FT), stl);
    }
    __syncwarp();
    // row r = (step r / 2, table r % 2) of the chunk: one bulk copy each, dealt to the warps (W_ = 2 T: one per warp)
#pragma unroll
    for (int r = warp; r < 2 * T; r += WA) {
      const int tt = r >> 1, tab = r & 1;
      const int st = __ldg(site_seq + min(max(pos + tt, 0), last));
      if (lane == 0) {
        const uint32_t dst = rows_addr + (uint32_t)(((b * T + tt) * 2 + tab) * (ROW * 4));
        lean_mbar_arrive_expect_tx(bar_full + 8 * b, (uint32_t)(ROW * sizeof(float)));
        lean_bulk_g2s(dst, ws + (size_t)(tab * N + st) * ROW, (uint32_t)(ROW * sizeof(float)), bar_full + 8 * b);
      }
    }
  };

  V Zr[NP], Zi[NP];  // z of this half's chain: units 32 p + hl (x) and 32 p + 16 + hl (y)
  float nlg;         // 32 - 2^26 (this lane's sum_j log2 |1 + z_j|^2): the constant term of the fixed-point addend
  unsigned n_acc = 0;

  // z <- z * row (in place); returns this lane's sum_j log2 |1 + z_j|^2.  rp: shared memory, lane offset applied.
  auto multiply = [&](const uint32_t rp, bool want_log) -> float {
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
            const float4 q = hw_lds128(rp + 16 * L * p);
            qlo[i] = make_float2(q.x, q.y);
            qhi[i] = make_float2(q.z, q.w);
          }
        }
#pragma unroll
        for (int i = 0; i < SG; ++i) {
          const int p = gq * GL + sq + i;
          if (sq + i < GL && p < NP) lean_cmul(Zr[p], Zi[p], qlo[i], qhi[i]);
        }
        if (want_log) {
#pragma unroll
          for (int i = 0; i < SG; ++i) {
            const int p = gq * GL + sq + i;
            if (sq + i < GL && p < NP) m[i] = add2_(Zr[p], one);
          }
#pragma unroll
          for (int i = 0; i < SG; ++i) {
            const int p = gq * GL + sq + i;
            if (sq + i < GL && p < NP) m[i] = P::mul(m[i], m[i]);
          }
#pragma unroll
          for (int i = 0; i < SG; ++i) {
            const int p = gq * GL + sq + i;
            if (sq + i < GL && p < NP) m[i] = P::fma(Zi[p], Zi[p], m[i]);
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

  // theta from scratch, then z = exp(-2 theta); padded units carry z = 0.  During the accumulation Zr[p] / Zi[p] hold
  // (Re, Im) theta of unit 32 p + hl / 32 p + 16 + hl -- the layout of the 8-byte loads of W, so the loop over the sites
  // is 2 NP loads and 2 NP packed FMAs (the (Re, Re) / (Im, Im) layout of z needed 144 register moves per site); only
  // the last pair can reach beyond M (M > 32 (NP - 1) by construction) and is the only one that tests its bounds.
  auto refresh = [&]() {
    const V *cb = reinterpret_cast<const V *>(net.b);
    const bool in0 = 2 * L * (NP - 1) + hl < M, in1 = 2 * L * (NP - 1) + L + hl < M;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int j0 = 2 * L * p + hl, j1 = j0 + L;
      Zr[p] = (cb && (p < NP - 1 || in0)) ? cb[j0] : P::mk(0.0f, 0.0f);
      Zi[p] = (cb && (p < NP - 1 || in1)) ? cb[j1] : P::mk(0.0f, 0.0f);
    }
    const V *wrow = reinterpret_cast<const V *>(net.W) + hl;
#pragma unroll 1
    for (int i = 0; i < N; ++i, wrow += M) {
      const float xv = ((sig_my_p[i >> 5] >> (i & 31)) & 1u) ? s1 : s0;
      const V xx = P::mk(xv, xv);
#pragma unroll
      for (int p = 0; p < NP - 1; ++p) {
        Zr[p] = P::fma(xx, __ldg(wrow + 2 * L * p), Zr[p]);
        Zi[p] = P::fma(xx, __ldg(wrow + 2 * L * p + L), Zi[p]);
      }
      if (in0) Zr[NP - 1] = P::fma(xx, __ldg(wrow + 2 * L * (NP - 1)), Zr[NP - 1]);
      if (in1) Zi[NP - 1] = P::fma(xx, __ldg(wrow + 2 * L * (NP - 1) + L), Zi[NP - 1]);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float e0 = exp_(-2.0f * Zr[p].x), e1 = exp_(-2.0f * Zi[p].x);
      float sn0, cs0, sn1, cs1;
      sincos_(-2.0f * Zr[p].y, &sn0, &cs0);
      sincos_(-2.0f * Zi[p].y, &sn1, &cs1);
      if (p == NP - 1 && !in0) e0 = 0.0f;
      if (p == NP - 1 && !in1) e1 = 0.0f;
      Zr[p] = P::mk(e0 * cs0, e1 * cs1);
      Zi[p] = P::mk(e0 * sn0, e1 * sn1);
    }
    float acc = 0.0f;
#pragma unroll
    for (int gq = 0; gq < NG; ++gq) {
      V prod = one;
#pragma unroll
      for (int p = gq * GL; p < (gq + 1) * GL && p < NP; ++p) {
        const V aa = add2_(Zr[p], one);
        V m = P::mul(aa, aa);
        m = P::fma(Zi[p], Zi[p], m);
        prod = P::mul(prod, m);
      }
      acc += P::lg2(prod.x * prod.y);
    }
    nlg = fmaf(acc, -LEAN_FX, 32.0f);
  };

  // one Metropolis step of both chains of the warp: table entry `e` (this step), rows at shared address `slot` (lane
  // offset applied), log2 u of this half's chain at tu_my + 4 tb.
  // A rejected move is undone by multiplying by the other row of the slot (the inverse of the one just applied).  The
  // undo RE-ENTERS the multiply code of the step (a loop that runs once, or twice for the half that rejected) instead
  // of being a second inlined copy behind a branch: with a copy, the two paths leave z in different registers and
  // ptxas pays ~27 MOVs on the ACCEPT path (95 % of the steps) to merge them; with the loop the new z always comes out
  // of the same instructions and only the rare back edge copies.  `undo` goes through an empty asm so that the
  // compiler cannot peel the loop back into the two copies.
  auto step_pf = [&](const int4 e, const uint32_t word, const int thr_u, const uint32_t slot) -> bool {
    const uint32_t msk = (uint32_t)e.y;
    const uint32_t waddr = tu_my + (uint32_t)e.x;
    const bool up = (word & msk) != 0u;
    uint32_t rp = slot + (up ? (uint32_t)(ROW * 4) : 0u);
    const int thr = thr_u - (up ? -e.z : e.z);
    int undo = 0;
    bool ok = false;
    for (;;) {
      const float acc = multiply(rp, true);
      asm volatile("" : "+r"(undo));
      if (undo) break;
      int v = __float2int_rn(fmaf(acc, LEAN_FX, nlg)) >> 6;
      if (hl == 0) v -= thr;
      const int s_lo = __reduce_add_sync(0xffffffffu, v & ~m_hi);
      const int s_hi = __reduce_add_sync(0xffffffffu, v & m_hi);
      if (__builtin_expect(((s_lo & ~m_hi) | (s_hi & m_hi)) > 0, 1)) {
        nlg = fmaf(acc, -LEAN_FX, 32.0f);
        ++n_acc;
        hw_sts32(waddr, word ^ msk);
        ok = true;
        break;
      }
      rp = 2u * slot + (uint32_t)(ROW * 4) - rp;
      undo = 1;
    }
    return ok;
  };
  auto step = [&](const int4 e, const uint32_t slot, const int tb) {
    const uint32_t msk = (uint32_t)e.y;
    const uint32_t waddr = tu_my + (uint32_t)e.x;
    const uint32_t word = hw_lds32(waddr);
    const bool up = (word & msk) != 0u;
    uint32_t rp = slot + (up ? (uint32_t)(ROW * 4) : 0u);
    const int thr = (int)hw_lds32(tu_my + 4u * (uint32_t)tb) - (up ? -e.z : e.z);
    int undo = 0;
    for (;;) {
      const float acc = multiply(rp, true);
      asm volatile("" : "+r"(undo));
      if (undo) break;
      // accept iff sum over the half's lanes of 2^20 (log2|1+z'|^2 - log2|1+z|^2) > 2^20 (log2 u - (+-A_site))
      int v = __float2int_rn(fmaf(acc, LEAN_FX, nlg)) >> 6;
      if (hl == 0) v -= thr;
      const int s_lo = __reduce_add_sync(0xffffffffu, v & ~m_hi);
      const int s_hi = __reduce_add_sync(0xffffffffu, v & m_hi);
      if (__builtin_expect(((s_lo & ~m_hi) | (s_hi & m_hi)) > 0, 1)) {
        nlg = fmaf(acc, -LEAN_FX, 32.0f);
        ++n_acc;
        hw_sts32(waddr, word ^ msk);
        break;
      }
      rp = 2u * slot + (uint32_t)(ROW * 4) - rp;
      undo = 1;
    }
  };

  // ---- prologue: rows of the first two chunks, random numbers of the first batch
  const uint32_t q0 = q0c;
  const uint32_t n_chunks = ((phase + n_steps + T - 1) >> LEAN_TSH) - q0;  // chunks that hold steps of this call
  issue_rows_tma(0, 0);
  if (n_chunks > 1) issue_rows_tma(1, 1);
  refill(0);
  const uint32_t lane_rows = rows_addr + 16u * (uint32_t)hl;
  const uint32_t tab_addr = lean_smem_u32(steptab);
  bool need_refresh = true;  // z is (re)computed from the configuration at the top of a chunk
  uint32_t s = 0;            // steps done
  int left = a.sweep_size;   // steps left in the current sweep
  int sweep = 0;
  auto emit = [&]() {
    if (sweep >= a.n_discard && lane < W32) {
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (c < n_valid)
          a.samples[((chain0 + c) * a.chain_length + (sweep - a.n_discard)) * W32 + lane] = sigx[c * CB + lane];
    }
  };
  // running buffer index and mbarrier parity of the current chunk (buf, par) and of the previous one (pbuf, ppar):
  // chunk j lives in buffer j % 3 on use (j / 3) of it
  uint32_t buf = 0, par = 0, pbuf = 2, ppar = 1;
  int pos = (int)(q0 << LEAN_TSH) - (int)phase;   // step index of the chunk's first slot (negative in the first chunk)
  for (uint32_t jq = 0; s < n_steps; ++jq, pos += T) {
    const int tb0 = (pos + (int)phase) & (SB - 1);
    if (tb0 == 0 && jq != 0) refill((uint32_t)(pos + (int)phase));
    if (need_refresh) {
      refresh();
      need_refresh = false;
    }
    hw_mbar_wait(bar_full + 8 * buf, par);  // rows and step table of this chunk
    const uint32_t cbase = lane_rows + (buf << LEAN_TSH) * (uint32_t)(2 * ROW * 4);
    const uint32_t tbase = tab_addr + (buf << LEAN_TSH) * 16u;
    const bool whole = pos >= 0 && pos + T <= (int)n_steps && left >= T;
    if (whole) {
      if constexpr (PF) {
        int4 e = hw_lds128i(tbase);
        uint32_t word = hw_lds32(tu_my + (uint32_t)e.x);
        int thr_u = (int)hw_lds32(tu_my + 4u * (uint32_t)tb0);
#pragma unroll UNR
        for (int tt = 0; tt < T; ++tt) {
          int4 en = e;
          uint32_t wn = word;
          int tn = thr_u;
          if (tt + 1 < T) {
            en = hw_lds128i(tbase + 16u * (tt + 1));
            wn = hw_lds32(tu_my + (uint32_t)en.x);
            tn = (int)hw_lds32(tu_my + 4u * (uint32_t)(tb0 + tt + 1));
          }
          const bool ok = step_pf(e, word, thr_u, cbase + (uint32_t)tt * (uint32_t)(2 * ROW * 4));
          if (tt + 1 < T && ok && en.x == e.x) wn ^= (uint32_t)e.y;  // the word was read before this step's flip
          e = en;
          word = wn;
          thr_u = tn;
        }
      } else {
#pragma unroll UNR
        for (int tt = 0; tt < T; ++tt) step(hw_lds128i(tbase + 16u * tt), cbase + (uint32_t)tt * (uint32_t)(2 * ROW * 4), tb0 + tt);
      }
      s += T;
      left -= T;
    } else {
      for (int tt = 0; tt < T; ++tt) {
        const int sidx = pos + tt;
        if (sidx < 0 || sidx >= (int)n_steps) continue;
        step(hw_lds128i(tbase + 16u * tt), cbase + (uint32_t)tt * (uint32_t)(2 * ROW * 4), tb0 + tt);
        ++s;
        if (--left == 0 && tt != T - 1) {
          // sweep boundary inside the chunk: emit the configurations
          left = a.sweep_size;
          __syncwarp();
          emit();
          ++sweep;
          if ((sweep % RATIO_REFRESH) == 0) need_refresh = true;
        }
      }
    }
    __syncwarp();
    if (left == 0) {
      left = a.sweep_size;
      emit();
      ++sweep;
      if ((sweep % RATIO_REFRESH) == 0) need_refresh = true;
    }
    // release this chunk's buffer; once every warp has released the PREVIOUS chunk's, request chunk j + 2 into it
    if (lane == 0) lean_mbar_arrive(bar_empty + 8 * buf);
    if (jq + 2 < n_chunks) {
      if (jq >= 1) hw_mbar_wait(bar_empty + 8 * pbuf, ppar);
      issue_rows_tma(jq + 2, pbuf);
    }
    pbuf = buf;
    ppar = par;
    if (++buf == 3u) {
      buf = 0;
      par ^= 1u;
    }
  }
  __syncwarp();
  if (lane < W32) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (c < n_valid) a.chain_state[(chain0 + c) * W32 + lane] = sigx[c * CB + lane];
  }
  if (a.n_accepted && hl == 0 && n_acc) atomicAdd(a.n_accepted, (unsigned long long)n_acc);
}

template <int NP, int GL, int UNR = 8, int W_ = 16, int SG = 2, int WA = W_, int PF = 0>
static int launch_ratio_lean_hw(nkfb_handle_t h, const nkfb_rbm_t *net, const SampArgs &a, const void *ws16,
                                const int *site_seq, const float *a0_seq, cudaStream_t st) {
  const size_t smem = lean_hw_smem_bytes(NP, W_, (net->n_visible + 31) >> 5);
  cudaError_t rc = cudaFuncSetAttribute(sample_ratio_lean_hw_kernel<NP, GL, UNR, W_, SG, WA, PF>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (rc != cudaSuccess) NKFB_FAIL(h, NKFB_ERR_UNSUPPORTED, "sampler needs %zu B of shared memory", smem);
  const int64_t blocks = (a.n_chains + (int64_t)2 * WA - 1) / ((int64_t)2 * WA);
  sample_ratio_lean_hw_kernel<NP, GL, UNR, W_, SG, WA, PF><<<(unsigned)blocks, W_ * 32, smem, st>>>(
      make_netdev<float>(net), (const float *)ws16, (float)a.s0, (float)a.s1, a, site_seq, a0_seq);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
