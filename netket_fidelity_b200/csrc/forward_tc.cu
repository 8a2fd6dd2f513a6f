// Tensor-core (tcgen05 + TMEM) version of pass A, the infidelity estimator, for fp32 parameters and
// NONE / GATE operators (BASELINE configs C1, C2, C4).
//
//   Theta[s, f] = sum_k X[s, k] * W~[k, f]      s: 128 sample pairs (UMMA M = 128)
//                                               k: spins + 1 bias row, padded to 16 (UMMA K = 16 per instruction)
//                                               f: NN <= 256 real columns (Re/Im interleaved) per chunk (UMMA N)
// X is exactly representable in bf16 (local states +-1 / 0,1); W~ is split into three bf16 terms
// W~ = hi + mid + lo (3 x 8 mantissa bits = fp32), so three UMMAs per k-step accumulate an fp32-grade
// product in TMEM.  Operands sit in shared memory in the canonical no-swizzle K-major layout
// [k / 8][row][8] (16-byte core-matrix rows); the accumulator is read back with tcgen05.ld, one
// thread per sample row, so the log-cosh row sums need no cross-thread reduction at all.
//
// log cosh z is accumulated in product form: sum_j (|x_j| - ln 2) + log prod_j (1 + t_j e^{-2 i y_j}),
// t_j = e^{-2|x_j|}, i.e. MUFU.EX2 + MUFU.SIN + MUFU.COS per hidden unit and one complex log per 8 units;
// the phase is therefore defined mod 2 pi (as any sum of principal-branch logs is, after exp()).
#include <cstdlib>

#include "tc_common.cuh"

namespace nkfb {

// ------------------------------------------------------------------ log-cosh accumulation in product form
// log cosh z = |x| + log((1 + t e^{-2iy}) / 2) (folded to Re z >= 0).  The factors (1 + t e^{-2iy}) / 2 (|.| <= 1, ~1 for
// small theta) are multiplied into ONE running complex product per evaluation, whose binary exponent is moved into the
// real sum every 8 units (two integer operations and two multiplies -- no logarithm, no arctangent inside the loop;
// round 1 took log + atan2 of the product every 8 units: 21 % of the kernel's instructions, profiles/r02_sass_hist_*).
// value() takes the one logarithm and the one arctangent at the end, so the phase is defined mod 2 pi as before.
struct LcAcc {
  float sx, sy, pr, pi;  // sum |x| + ln 2 * (exponents taken out of the product), sum of the folded y, running product
  __device__ __forceinline__ void reset() {
    sx = sy = pi = 0.f;
    pr = 1.f;
  }
  __device__ __forceinline__ void add(float x, float y) {
    // fold theta to Re >= 0: (x, y) -> (|x|, y sgn x), the sign bit of x moved onto y with one logic operation
    y = __uint_as_float(__float_as_uint(y) ^ (__float_as_uint(x) & 0x80000000u));
    x = fabsf(x);
    const float t = mufu_ex2(x * (float)(-2.0 * LOG2E));
    const float a = 2.f * y;
    float c, s;
#ifdef NKFB_TC_POLY_SINCOS
    if (fabsf(a) <= 1.f) {
      const float a2 = a * a;
      s = a * fmaf(a2, fmaf(a2, fmaf(a2, fmaf(a2, 2.7557319e-6f, -1.9841270e-4f), 8.3333333e-3f), -1.6666667e-1f), 1.f);
      c = fmaf(a2, fmaf(a2, fmaf(a2, fmaf(a2, fmaf(a2, -2.7557319e-7f, 2.4801587e-5f), -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.f);
    } else {
      sincosf(a, &s, &c);
    }
#else
    c = mufu_cos(a);
    s = mufu_sin(a);
#endif
    const float ht = 0.5f * t;
    const float wr = fmaf(ht, c, 0.5f), wi = -ht * s;
    const float nr = pr * wr - pi * wi, ni = fmaf(pr, wi, pi * wr);
    pr = nr;
    pi = ni;
    sx += x;
    sy += y;
  }
  // product <- product * 2^-e, sx <- sx + e ln 2, e = binary exponent of max(|pr|, |pi|): the product stays in [1, 2) x
  // [0, 2) whatever the number of units (a zero product stays zero and ends as log 0 = -inf, like the log before)
  __device__ __forceinline__ void renorm() {
    const int eb = __float_as_int(fmaxf(fabsf(pr), fabsf(pi))) & 0x7F800000;
    const float sc = __int_as_float(0x7F000000 - eb);
    pr *= sc;
    pi *= sc;
    sx = fmaf((float)((eb >> 23) - 127), 0.69314718055994530942f, sx);
  }
  __device__ __forceinline__ cx<float> value() const {
    return mk<float>(sx + 0.5f * logf(fmaf(pr, pr, pi * pi)), sy + atan2f(pi, pr));
  }
};

// Epilogue of one (chunk, tile, side): this thread's row, columns [c_begin, c_end) of the chunk (multiples of 8).
// FLIP: also accumulate the configuration with the gate site flipped (theta + d W[idx, :]).  Columns beyond 2M (last
// chunk) need no test: W~ and the staged gate row are zero there, so theta = theta' = 0 and the factor is exactly 1.
template <bool FLIP>
__device__ __forceinline__ void fwd_epilogue(LcAcc &A0, LcAcc &A1, uint32_t trow, int c_begin, int c_end, const float *wg,
                                             float d) {
  int c0 = c_begin;
#pragma unroll 1
  for (; c0 + 16 <= c_end; c0 += 16) {
    float v[16];
    tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      A0.add(v[2 * u], v[2 * u + 1]);
      if (FLIP) A1.add(fmaf(d, wg[c0 + 2 * u], v[2 * u]), fmaf(d, wg[c0 + 2 * u + 1], v[2 * u + 1]));
    }
    A0.renorm();
    if (FLIP) A1.renorm();
  }
  if (c0 < c_end) {  // a last block of 8 columns
    float v[8];
    tmem_ld8(trow + (uint32_t)c0, v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      A0.add(v[2 * u], v[2 * u + 1]);
      if (FLIP) A1.add(fmaf(d, wg[c0 + 2 * u], v[2 * u]), fmaf(d, wg[c0 + 2 * u + 1], v[2 * u + 1]));
    }
    A0.renorm();
    if (FLIP) A1.renorm();
  }
}

struct TcFwdArgs {
  OpDev op1, op2, op4;
  const uint32_t *sigma, *eta;
  int64_t S;
  float s0, s1;
  float cv;
  int has_cv;
  cx<float> *log_val;
  float *L;
  cx<float> *rho1;
  double *sums;
  const __nv_bfloat16 *prep_psi, *prep_phi;
  TcGeom g;
  int64_t n_groups;
  int tmem_cols;  // TMEM columns this CTA allocates
};

__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes)
               : "memory");
}
// TMA bulk copy global -> shared (SASS UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

// FM: bit e - 1 set <=> evaluation e (1 = (psi, eta) with U2, 2 = (phi, sigma) with U1, 3 = (phi, eta) with U4) goes
// through a gate, i.e. has a second (flipped) connected element.  A template parameter so that each of the 16
// unrolled (network, tile, side) phases carries exactly one epilogue body (the kernel was 51 000 SASS instructions
// with both bodies and the log / atan2 flushes inlined; `no_instruction` stalls 0.65 per issued instruction).
//
// Roles (round 2): warps 0..7 build the X tiles and run the log-cosh epilogues; warp 8 issues the TMA bulk copies of
// the W~ chunks and the UMMAs.  Phases q = 0, 1, 2, ... run over (group, network, chunk, tile, side); the hand-offs are
// mbarriers, there is no CTA barrier in the loop:
//   mbar_x  (8 arrivals, one per epilogue warp)  X tile of phase q written (and, before a chunk's first phase, its gate
//           rows); its completion also tells the issuer that every warp is past epilogue q - 2: TMEM buffer q & 1 is free
//   mbar_b  (bytes)   W~ chunk landed                       mbar_a  (tcgen05.commit)  UMMAs of phase q complete
// An epilogue warp in phase q: wait a(q) -> X tile q + 1 -> arrive x -> epilogue q, so the tensor pipe computes phase
// q + 1 under the epilogue of phase q (two accumulators in TMEM).  With thread 0 of warp 0 issuing the UMMAs (21 per
// phase plus their descriptors) that warp reached every CTA barrier late: `barrier` was the top stall (1.6 per issued
// instruction, ncu profiles/r02_*).
constexpr int TC_FWD_THREADS = TC_THREADS + 32;

template <int FM>
__global__ void __launch_bounds__(TC_FWD_THREADS, 2) infid_fwd_tc_kernel(NetDev<float> psi, NetDev<float> phi, TcFwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcGeom g = a.g;
  unsigned char *Bs = smem;                             // 3 splits of the current chunk (written by TMA bulk copies only)
  unsigned char *As = Bs + g.b_chunk_bytes();           // X tile
  uint4 *xtab = reinterpret_cast<uint4 *>(As + g.a_bytes());  // 256 x 16 B: bf16 images of 8 spins (tc_build_x_tile_fast)
  float *wg = reinterpret_cast<float *>(xtab + 256);        // [2 chunk parities][2 sides][NN] floats: row `idx` of W
  cx<float> *xch = reinterpret_cast<cx<float> *>(wg + 4 * g.NN);  // exchange buffer [7][TC_ROWS]
  __shared__ __align__(8) uint64_t mbar, mbar_b, mbar_x;
  __shared__ uint32_t tmem_base_sh;
  __shared__ double red[16];
  __shared__ __align__(16) uint4 xtail[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane;  // TMEM lane == sample row of the tile
  const int half = (warp >> 2) & 1;        // which half of the chunk's columns this thread reduces
  const bool issuer = warp == TC_THREADS / 32;
  const uint32_t mbar_a = smem_u32(&mbar), mbar_ba = smem_u32(&mbar_b), mbar_xa = smem_u32(&mbar_x);
  const uint32_t b_bytes = (uint32_t)g.b_chunk_bytes();

  if (tid == 0) {
    mbar_init(mbar_a, 1);
    mbar_init(mbar_ba, 1);
    mbar_init(mbar_xa, TC_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // 256 columns per CTA (two CTAs share the SM's 512); all 512 when the CTA is alone on its SM (a.tmem_cols)
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_sh), a.tmem_cols);
  tc_fill_x_table(xtab, a.s0, a.s1);
  tc_fill_x_tail(xtail, g.N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;
  const uint32_t idesc = umma_idesc_bf16_f32(TC_ROWS, g.NN, 0, 0);

  const float ssum = a.s0 + a.s1;
  const int W32 = (g.N + 31) >> 5;
  const int cols = 2 * g.M;
  const bool dbuf = 2 * g.NN <= a.tmem_cols;
  double sL = 0, sL2 = 0, sA2 = 0, sRe = 0, sn = 0, sbad = 0;  // sbad: local estimators that are NaN or infinite

  if (issuer) {
    // ------------------------------------------------------------------------------------------ the issuing warp
    uint32_t q = 0, cc = 0;  // phase counter, chunk counter
    if (lane == 0 && (int64_t)blockIdx.x < a.n_groups) {
      tc_mbar_expect_tx(mbar_ba, b_bytes);
      tc_bulk_g2s(smem_u32(Bs), a.prep_psi, b_bytes, mbar_ba);
    }
    for (int64_t grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
      for (int netid = 0; netid < 2; ++netid) {
        for (int chunk = 0; chunk < g.NC; ++chunk, ++cc) {
          for (int p = 0; p < 2 * TC_T; ++p, ++q) {
            mbar_wait(mbar_xa, q & 1);  // X tile of phase q written; every warp is past epilogue q - 2
            if (p == 0) mbar_wait_relaxed(mbar_ba, cc & 1);  // this chunk's W~
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs);
              const uint32_t td = tmem_d + (dbuf ? (uint32_t)((q & 1) * g.NN) : 0u);
              for (int s = 0; s < 3; ++s)
                for (int ks = 0; ks < g.KS; ++ks) {
                  const uint64_t ad = umma_desc(a_base + ks * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128);
                  const uint64_t bd = umma_desc(b_base + s * (uint32_t)g.b_split_bytes() + ks * 2 * (g.NN * 16),
                                                g.NN * 16, 128);
                  umma_bf16(td, ad, bd, idesc, (s | ks) ? 1u : 0u);
                }
              umma_commit(mbar_a);
            }
            __syncwarp();
            if (p == 2 * TC_T - 1) {
              // every UMMA of this chunk must complete before its W~ buffer is overwritten by the next chunk (of this
              // network, of the other one, or the first of the CTA's next group): requested now, lands under the epilogue
              mbar_wait_relaxed(mbar_a, q & 1);
              const bool last_chunk = chunk + 1 == g.NC;
              const bool more = !(last_chunk && netid == 1) || grp + gridDim.x < a.n_groups;
              if (more && lane == 0) {
                const __nv_bfloat16 *nprep = (last_chunk ? 1 - netid : netid) == 0 ? a.prep_psi : a.prep_phi;
                const int nchunk = last_chunk ? 0 : chunk + 1;
                tc_mbar_expect_tx(mbar_ba, b_bytes);
                tc_bulk_g2s(smem_u32(Bs), reinterpret_cast<const unsigned char *>(nprep) + (size_t)nchunk * b_bytes, b_bytes,
                            mbar_ba);
              }
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ the epilogue warps
    uint32_t q = 0, cc = 0;
    // this warp's writes to shared memory are done and visible to the tensor pipe: one arrive per warp
    auto warp_arrive_x = [&]() {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(mbar_xa) : "memory");
    };
    // row `idx` of W (columns of chunk `chunk_n` of network `netid_n`) for the evaluations that flip a spin
    auto stage_wg = [&](int netid_n, int chunk_n, uint32_t cc_n) {
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int ev = netid_n * 2 + side;
        if (ev != 0 && ((FM >> (ev - 1)) & 1)) {
          const OpDev &op = (ev == 1) ? a.op2 : (ev == 2 ? a.op1 : a.op4);
          const float *Wn = netid_n == 0 ? psi.W : phi.W;
          float *dst = wg + ((cc_n & 1) * 2 + side) * g.NN;
          for (int n = tid; n < g.NN; n += TC_THREADS) {
            const int col = chunk_n * g.NN + n;
            dst[n] = col < cols ? Wn[(int64_t)op.idx * cols + col] : 0.f;
          }
        }
      }
    };
    auto build_x = [&](int64_t grp_n, int p_n) {
      const int t = p_n >> 1, side = p_n & 1;
      const uint32_t *packed = side == 0 ? a.sigma : a.eta;
      const int64_t s_base = (grp_n * TC_T + t) * TC_ROWS;
      if (g.KC <= 16)
        tc_build_x_tile_fast<TC_THREADS / TC_ROWS>(As, packed, s_base, a.S, g, xtab, xtail);
      else
        tc_build_x_tile(As, packed, s_base, a.S, g, a.s0, a.s1, false, xtab);
    };
    if ((int64_t)blockIdx.x < a.n_groups) {  // prologue: gate rows of the first chunk, X tile of phase 0
      stage_wg(0, 0, 0);
      build_x(blockIdx.x, 0);
      warp_arrive_x();
    }

    for (int64_t grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
      // accumulators: [tile][eval] with eval 0 = (psi,sigma) 1 = (psi,eta) 2 = (phi,sigma) 3 = (phi,eta); flips for 1,2,3
      LcAcc acc0[TC_T][4], acc1[TC_T][4];
#pragma unroll
      for (int t = 0; t < TC_T; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc0[t][e].reset();
          acc1[t][e].reset();
        }

#pragma unroll
      for (int netid = 0; netid < 2; ++netid) {
#pragma unroll 1
        for (int chunk = 0; chunk < g.NC; ++chunk, ++cc) {
          // columns per half (multiple of 8); the last chunk stops at the last real column (rounded up to 16)
          const int nh = min(g.NN, ((cols - chunk * g.NN + 15) >> 4) << 4) / 2;
#pragma unroll
          for (int p = 0; p < 2 * TC_T; ++p, ++q) {
            const int t = p >> 1, side = p & 1;
            const int ev = netid * 2 + side;
            const OpDev &op = (ev == 1) ? a.op2 : (ev == 2 ? a.op1 : a.op4);
            const bool flip = ev != 0 && ((FM >> (ev == 0 ? 0 : ev - 1)) & 1);
            const uint32_t *packed = side == 0 ? a.sigma : a.eta;
            const int64_t s_base = (grp * TC_T + t) * TC_ROWS;
            // what phase q + 1 is: next (tile, side) of this chunk, the first of the next chunk / network, or of the next group
            const bool last_chunk = chunk + 1 == g.NC;
            const bool more = p + 1 < 2 * TC_T || !(last_chunk && netid == 1) || grp + gridDim.x < a.n_groups;
            auto next_phase = [&]() {
              if (!more) return;
              if (p + 1 < 2 * TC_T) {
                build_x(grp, p + 1);
              } else {
                stage_wg(last_chunk ? 1 - netid : netid, last_chunk ? 0 : chunk + 1, cc + 1);
                build_x((last_chunk && netid == 1) ? grp + gridDim.x : grp, 0);
              }
              warp_arrive_x();
            };
            mbar_wait(mbar_a, q & 1);  // UMMAs of phase q complete: the X buffer is free, the accumulator is ready
            tc_fence_after();
            if (dbuf) next_phase();
            // ---- epilogue: this thread's row, its half of the chunk's columns
            float d = 0.f;
            if (flip) {
              const int64_t gs = s_base + row;
              if (gs < a.S) {
                const uint32_t w = packed[gs * W32 + (op.idx >> 5)];
                d = ssum - 2.f * (((w >> (op.idx & 31)) & 1u) ? a.s1 : a.s0);
              }
            }
            LcAcc &A0 = acc0[t][ev], &A1 = acc1[t][ev];
            const uint32_t trow = tmem_d + (dbuf ? (uint32_t)((q & 1) * g.NN) : 0u) + ((uint32_t)(32 * (warp & 3)) << 16);
            const float *wgp = wg + ((cc & 1) * 2 + side) * g.NN;
            if (flip)
              fwd_epilogue<true>(A0, A1, trow, half * nh, (half + 1) * nh, wgp, d);
            else
              fwd_epilogue<false>(A0, A1, trow, half * nh, (half + 1) * nh, wgp, d);
            if (!dbuf) next_phase();  // one accumulator: the next UMMAs may only start after this epilogue
          }
        }
      }

      // ---- combine the two column halves, add the visible-bias terms, finish the estimator
#pragma unroll
      for (int t = 0; t < TC_T; ++t) {
        const int64_t s_base = (grp * TC_T + t) * TC_ROWS;
        asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory");  // the epilogue warps only
        if (half == 1) {
          xch[0 * TC_ROWS + row] = acc0[t][0].value();
          xch[1 * TC_ROWS + row] = acc0[t][1].value();
          xch[2 * TC_ROWS + row] = acc0[t][2].value();
          xch[3 * TC_ROWS + row] = acc0[t][3].value();
          if (FM & 1) xch[4 * TC_ROWS + row] = acc1[t][1].value();
          if (FM & 2) xch[5 * TC_ROWS + row] = acc1[t][2].value();
          if (FM & 4) xch[6 * TC_ROWS + row] = acc1[t][3].value();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory");
        const int64_t gs = s_base + row;
        if (half == 0 && gs < a.S) {
          cx<float> b0[4], b1[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) b0[e] = acc0[t][e].value() + xch[e * TC_ROWS + row];
          b1[0] = mk<float>(0, 0);
#pragma unroll
          for (int e = 1; e < 4; ++e)
            b1[e] = ((FM >> (e - 1)) & 1) ? acc1[t][e].value() + xch[(3 + e) * TC_ROWS + row] : mk<float>(0, 0);
          // visible-bias terms a.x for (psi, sigma) (psi, eta) (phi, sigma) (phi, eta)
          cx<float> ax[4] = {mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0)};
          if (psi.a || phi.a) {
            for (int i = 0; i < g.N; ++i) {
              const float xs = ((a.sigma[gs * W32 + (i >> 5)] >> (i & 31)) & 1u) ? a.s1 : a.s0;
              const float xe = ((a.eta[gs * W32 + (i >> 5)] >> (i & 31)) & 1u) ? a.s1 : a.s0;
              if (psi.a) {
                const cx<float> ai = mk<float>(psi.a[2 * i], psi.a[2 * i + 1]);
                ax[0] = ax[0] + xs * ai;
                ax[1] = ax[1] + xe * ai;
              }
              if (phi.a) {
                const cx<float> ai = mk<float>(phi.a[2 * i], phi.a[2 * i + 1]);
                ax[2] = ax[2] + xs * ai;
                ax[3] = ax[3] + xe * ai;
              }
            }
          }
          cx<float> T[4], rho = mk<float>(0, 0);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const OpDev &op = (e == 1) ? a.op2 : (e == 2 ? a.op1 : a.op4);
            const NetDev<float> &net = e < 2 ? psi : phi;
            const uint32_t *packed = (e & 1) ? a.eta : a.sigma;
            const cx<float> l0 = b0[e] + ax[e];
            if (e == 0 || !((FM >> (e == 0 ? 0 : e - 1)) & 1)) {
              T[e] = l0;
            } else {
              const int bit = (packed[gs * W32 + (op.idx >> 5)] >> (op.idx & 31)) & 1u;
              const float dd = ssum - 2.f * (bit ? a.s1 : a.s0);
              cx<float> l1 = b1[e] + ax[e];
              if (net.a) l1 = l1 + dd * mk<float>(net.a[2 * op.idx], net.a[2 * op.idx + 1]);
              const cx<float> m0 = mk<float>((float)op.gate_mel[bit][0][0], (float)op.gate_mel[bit][0][1]);
              const cx<float> m1 = mk<float>((float)op.gate_mel[bit][1][0], (float)op.gate_mel[bit][1][1]);
              const float amax = fmaxf(l0.re, l1.re);
              const cx<float> e0 = m0 * cexp(mk<float>(l0.re - amax, l0.im));
              const cx<float> e1 = m1 * cexp(mk<float>(l1.re - amax, l1.im));
              const cx<float> sum = e0 + e1;
              T[e] = clog(sum);
              T[e].re += amax;
              if (e == 1) rho = (1.f / abs2(sum)) * (e1 * conj(sum));
            }
          }
          // l = T(phi,sigma) + T(psi,eta) - T(psi,sigma) - T(phi,eta)
          const cx<float> l = T[2] + T[1] - T[0] - T[3];
          const cx<float> A = cexp(l);
          const float A2 = abs2(A);
          float Lv = A.re;
          if (a.has_cv) Lv += a.cv * (A2 - 1.f);
          a.log_val[gs] = l;
          a.L[gs] = Lv;
          a.rho1[gs] = rho;
          sL += (double)Lv;
          sL2 += (double)Lv * (double)Lv;
          sA2 += (double)A2;
          sRe += (double)A.re;
          sn += 1.0;
          if (!isfinite(Lv)) sbad += 1.0;
        }
      }
    }
  }

  double v[6] = {sL, sL2, sA2, sRe, sn, sbad};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double r = block_sum_256(v[q], red);  // all 9 warps (the issuing warp adds zeros)
    if (threadIdx.x == 0) atomicAdd(a.sums + q, r);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, a.tmem_cols);
}

// geometry of the tensor-core estimator for a pair of networks; false when the tensor-core path does not apply to them
static bool tc_fwd_geometry(const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, TcGeom *g, size_t *smem_out) {
  if (psi->dtype != NKFB_F32 || phi->dtype != NKFB_F32) return false;
  if (psi->n_hidden != phi->n_hidden || psi->n_visible != phi->n_visible) return false;
  // two CTAs per SM (each its own TMEM accumulator): one CTA's UMMAs and X-tile builds overlap the other's
  // log-cosh epilogue; the W~ chunk of a CTA is sized for half of the shared memory
  size_t budget = 104 * 1024;
  if (const char *e = getenv("NKFB_TC_FWD_BUDGET_KB")) budget = (size_t)atoi(e) * 1024;
  const size_t xb = (size_t)(((psi->n_visible + 1 + 15) / 16) * 2) * TC_ROWS * 16 + 4096;  // X tile + spin-group table
  if (!tc_geometry(psi->n_visible, psi->n_hidden, g, 2 * psi->n_hidden, 0, xb, budget) &&
      !tc_geometry(psi->n_visible, psi->n_hidden, g, 2 * psi->n_hidden, 0, xb))
    return false;
  const size_t smem = g->b_chunk_bytes() + g->a_bytes() + 4096 + (size_t)g->NN * 16 + 7 * TC_ROWS * sizeof(cx<float>) + 128;
  if (smem > 227 * 1024) return false;
  *smem_out = smem;
  return true;
}

static nkfb_handle_s::TcTablesKey tc_fwd_key_of(const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const TcGeom &g) {
  nkfb_handle_s::TcTablesKey k;
  k.w[0] = psi->kernel;
  k.w[1] = phi->kernel;
  k.b[0] = psi->bias;
  k.b[1] = phi->bias;
  k.n_visible = psi->n_visible;
  k.n_hidden = psi->n_hidden;
  k.NN = g.NN;
  k.NC = g.NC;
  k.KC = g.KC;
  return k;
}

// The bf16 splits of both networks (the handle's second workspace).  With NKFB_OPT_TC_TABLES_VALID set and the tables of
// these very parameter buffers and geometry already there (nkfb_tc_prepare, or an earlier estimator launch), nothing is
// launched: `force` = false is the estimator's own call, true is nkfb_tc_prepare's.
static int tc_fwd_tables(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const TcGeom &g, bool force,
                         cudaStream_t st) {
  const size_t need = 2 * g.prep_bytes_per_net();
  if (h->ws2_bytes < need) {
    if (h->ws2) cudaFree(h->ws2);
    h->ws2 = nullptr;
    h->ws2_bytes = 0;
    h->tc_fwd_key.w[0] = nullptr;
    NKFB_CHECK_CUDA(h, cudaMalloc(&h->ws2, need));
    h->ws2_bytes = need;
  }
  const nkfb_handle_s::TcTablesKey key = tc_fwd_key_of(psi, phi, g);
  if (!force && h->opt_tc_tables_valid && tc_key_equal(h->tc_fwd_key, key)) return NKFB_OK;
  __nv_bfloat16 *prep_psi = (__nv_bfloat16 *)h->ws2;
  __nv_bfloat16 *prep_phi = (__nv_bfloat16 *)((char *)h->ws2 + g.prep_bytes_per_net());
  const int64_t tot = (int64_t)g.NC * g.KC * g.NN;
  const int grid = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 8);
  tc_prep_kernel<<<grid, 256, 0, st>>>(make_netdev<float>(psi), g, prep_psi);
  NKFB_LAUNCHED(h);
  tc_prep_kernel<<<grid, 256, 0, st>>>(make_netdev<float>(phi), g, prep_phi);
  NKFB_LAUNCHED(h);
  h->tc_fwd_key = key;
  return NKFB_OK;
}

// nkfb_tc_prepare, estimator half: returns NKFB_OK (also when the tensor-core path does not apply: nothing to prepare)
int tc_fwd_prepare(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const double ls[2], cudaStream_t st) {
  TcGeom g;
  size_t smem;
  if (!bf16_exact(ls[0]) || !bf16_exact(ls[1]) || !tc_fwd_geometry(psi, phi, &g, &smem)) return NKFB_OK;
  return tc_fwd_tables(h, psi, phi, g, true, st);
}

// returns NKFB_OK, an error, or 1 when the tensor-core path does not apply (caller falls back to the SIMT kernel)
int tc_infidelity_fwd(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_rbm_t *phi, const nkfb_op_t *op1,
                      const nkfb_op_t *op2, const nkfb_op_t *op4, const uint32_t *sigma, const uint32_t *eta,
                      int64_t S, const double ls[2], double cv, void *log_val, void *L, void *rho1, double *sums,
                      cudaStream_t st) {
  if (psi->dtype != NKFB_F32) return 1;
  const nkfb_op_t *ops[3] = {op1, op2, op4};
  for (int i = 0; i < 3; ++i)
    if (ops[i] && ops[i]->kind == NKFB_OP_ISING) return 1;
  if (psi->n_hidden != phi->n_hidden) return 1;
  if (!bf16_exact(ls[0]) || !bf16_exact(ls[1])) return 1;
  if (S < 2048) return 1;
  TcGeom g;
  size_t smem;
  if (!tc_fwd_geometry(psi, phi, &g, &smem)) return 1;
  {
    const int rc = tc_fwd_tables(h, psi, phi, g, false, st);
    if (rc != NKFB_OK) return rc;
  }
  __nv_bfloat16 *prep_psi = (__nv_bfloat16 *)h->ws2;
  __nv_bfloat16 *prep_phi = (__nv_bfloat16 *)((char *)h->ws2 + g.prep_bytes_per_net());
  TcFwdArgs a;
  a.op1 = make_opdev(op1);
  a.op2 = make_opdev(op2);
  a.op4 = make_opdev(op4);
  a.sigma = sigma;
  a.eta = eta;
  a.S = S;
  a.s0 = (float)ls[0];
  a.s1 = (float)ls[1];
  a.has_cv = !isnan(cv);
  a.cv = a.has_cv ? (float)cv : 0.f;
  a.log_val = (cx<float> *)log_val;
  a.L = (float *)L;
  a.rho1 = (cx<float> *)rho1;
  a.sums = sums;
  a.prep_psi = prep_psi;
  a.prep_phi = prep_phi;
  a.g = g;
  const int64_t tiles = (S + TC_ROWS - 1) / TC_ROWS;
  a.n_groups = (tiles + TC_T - 1) / TC_T;
  a.tmem_cols = smem > (size_t)115 * 1024 ? 512 : 256;  // more than half of the SM's shared memory: one CTA per SM
  const int per_sm = smem <= 112 * 1024 ? 2 : 1;
  int grid = (int)std::min<int64_t>(a.n_groups, (int64_t)h->sm_count * per_sm);
  if (h->opt_fwd_max_ctas > 0) grid = std::min(grid, h->opt_fwd_max_ctas);  // NKFB_OPT_FWD_MAX_CTAS
  // which evaluations go through a gate: bit 0 (psi, eta) U2, bit 1 (phi, sigma) U1, bit 2 (phi, eta) U4
  const int fm = (a.op2.kind == NKFB_OP_GATE ? 1 : 0) | (a.op1.kind == NKFB_OP_GATE ? 2 : 0) | (a.op4.kind == NKFB_OP_GATE ? 4 : 0);
#define NKFB_FWD_TC_CASE(FM)                                                                                              \
  case FM:                                                                                                                \
    NKFB_CHECK_CUDA(h, cudaFuncSetAttribute(infid_fwd_tc_kernel<FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    infid_fwd_tc_kernel<FM><<<grid, TC_FWD_THREADS, smem, st>>>(make_netdev<float>(psi), make_netdev<float>(phi), a);         \
    break;
  switch (fm) {
    NKFB_FWD_TC_CASE(0) NKFB_FWD_TC_CASE(1) NKFB_FWD_TC_CASE(2) NKFB_FWD_TC_CASE(3)
    NKFB_FWD_TC_CASE(4) NKFB_FWD_TC_CASE(5) NKFB_FWD_TC_CASE(6) NKFB_FWD_TC_CASE(7)
  }
#undef NKFB_FWD_TC_CASE
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
