// Tensor-core (tcgen05 + TMEM) version of pass B, the score-function gradient, for fp32 parameters and
// NONE / GATE operators.  A CTA owns one chunk of NN real columns of the (N+1) x 2(M+1) gradient tile
// [rows: spins + bias row; columns: hidden units Re/Im + one pseudo-unit for the visible bias] and a range of
// sample tiles.  Per tile and side (sigma, eta):
//   GEMM 1  Theta[s, f]  = X[s, :] W~[:, f]            UMMA 128 x NN x 16, A = X tile (K-major), bf16 x 3 splits
//   epilogue             Y[s, f] = w_s conj(tanh Theta[s, f])  (+ the flipped connected element for a gate)
//                        one thread per sample row, written back to shared memory as three bf16 splits
//   GEMM 2  G[i, f]     += X[s, i] Y[s, f]             UMMA 128 x NN x 16, A = the SAME X tile read MN-major
//                                                      (rows = spins), B = Y MN-major; G stays in TMEM
// and at the end G is read from TMEM and added to the fp64 accumulators with one atomic per element.
#include <cstdlib>

#include <type_traits>

#include "tc_common.cuh"

namespace nkfb {

constexpr int GX_CHUNKS = 16;  // the X tile always spans 128 "spin rows" (UMMA M of GEMM 2), zero padded

struct TcGradArgs {
  OpDev op2;
  const uint32_t *sigma, *eta;
  int64_t S;
  float s0, s1, cv;
  int has_cv;
  const cx<float> *log_val, *rho1;
  const double *sums;
  double *grad_acc;
  const __nv_bfloat16 *prep;
  TcGeom g;
  int64_t tiles_per_split;
  int has_bias, has_vbias;
  const cx<float> *weights;  // weighted-score mode (nkfb_weighted_score): w_sig = weights[s], sigma side only
};

// tanh z with MUFU: sg ((1 - t^2) + 2 i t sin 2y) / (1 + 2 t cos 2y + t^2), t = e^{-2|x|}
// (tanh is odd and the imaginary part 2 t sin 2y / den is even in the fold (x, y) -> (-x, -y), so only the sign of
// the real part has to follow x: no selects)
__device__ __forceinline__ cx<float> ctanh_fast(float x, float y) {
  const float t = mufu_ex2(fabsf(x) * (float)(-2.0 * LOG2E));
  const float a = 2.f * y;
  const float c = mufu_cos(a), s = mufu_sin(a);
  float inv;  // 1 / (1 + 2 t cos 2y + t^2): one MUFU.RCP (__fdividef carries a range-check slow path: 14 instructions here)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaf(t, fmaf(2.f, c, t), 1.f)));
  return mk<float>(copysignf((1.f - t * t) * inv, x), 2.f * t * s * inv);
}

__device__ __forceinline__ void split3(float a, float b, uint32_t &hi, uint32_t &mid, uint32_t &lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float ra = a - __low2float(h), rb = b - __high2float(h);
  const __nv_bfloat162 m = __floats2bfloat162_rn(ra, rb);
  const float qa = ra - __low2float(m), qb = rb - __high2float(m);
  const __nv_bfloat162 l = __floats2bfloat162_rn(qa, qb);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  mid = *reinterpret_cast<const uint32_t *>(&m);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// NB8: 8-column blocks of the chunk per thread.  The correction of the gate row, sum_s d_s w_eta conj(rho_1)
// conj(tanh theta'), is accumulated per thread in 8 NB8 registers over ALL its sample tiles and reduced over the 128
// rows once at the end (a per-tile reduction costs 40 shuffles + 8 shared atomics per 8 columns).
// NT threads: 256 (two column halves per sample row) or 512 (four quarters: sixteen warps hide the latencies of the
// thread-per-row epilogue twice as well).
//
// Software pipeline over the (tile, side) iterations k of a CTA (round 2; round 1 ran GEMM 1 -> epilogue -> GEMM 2 ->
// X tile strictly one after the other and spent 42 % of its stall samples spinning on the two mbarriers):
//   two X-tile buffers, two Theta accumulators in TMEM (+ the resident G), one Y buffer;  iteration k:
//     wait GEMM 1(k)                                       Theta[k & 1] ready (issued during iteration k - 1)
//     epilogue block 0 (registers only)                    overlaps GEMM 2(k - 1)
//     wait GEMM 2(k - 1)                                   Y buffer and X[(k + 1) & 1] are free
//     store block 0;  weights of the next tile;  X tile k + 1;  barrier;  issue GEMM 1(k + 1)
//     epilogue blocks 1 ..                                 overlaps GEMM 1(k + 1)
//     barrier;  issue GEMM 2(k)
// so the tensor pipe always works under an epilogue.  The UMMAs are issued by a DEDICATED warp (warp NT / 32; the CTA
// has NT + 32 threads): with thread 0 of an epilogue warp issuing them (45 UMMAs and their descriptors per iteration) that
// warp reached every CTA barrier late and `barrier` became the top stall (2.2 per issued instruction, ncu).  There is no
// CTA barrier in the loop: the epilogue warps signal "X tile written" and "Y written" on two mbarriers (one arrive per
// warp) that only the issuing warp waits on.  TAIL = the chunk holds the visible-bias pseudo-unit or padding
// (last chunk only): the per-unit index tests exist in that instantiation alone.
// Two-level accumulation of G (round 2).  The tensor pipe aligns every addend to the accumulator and TRUNCATES it, so a
// long chain of UMMAs into one fp32 accumulator drifts by ~0.05 eps per sample relative to a non-zero mean (measured
// against the complex128 kernels on the same 2^20 samples: 2.0e-4 of max |g| after 65 536 samples per CTA, 1.2e-5 after
// 4 096; the CUDA-core kernel: 2.0e-6).  G is therefore accumulated by the UMMAs over TC_G_FLUSH iterations only
// (4 tiles of both sides = 512 samples) in a first TMEM accumulator, which the epilogue warps then add -- rounded to
// nearest, 60 instructions per thread per 8 iterations -- into a second, resident one.
constexpr int TC_G_FLUSH = 8;

struct GradEpi {
  uint32_t hi[4], mid[4], lo[4];
};

template <int NB8, int NT>
__global__ void __launch_bounds__(NT + 32, 1) infid_grad_tc_kernel(NetDev<float> psi, TcGradArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcGeom g = a.g;
  const int NN = g.NN;
  constexpr size_t X_BYTES = (size_t)GX_CHUNKS * TC_ROWS * 16;
  const size_t y_split_bytes = (size_t)(NN / 8) * TC_ROWS * 16;
  unsigned char *Bs = smem;                                      // W~ chunk, 3 splits, resident
  unsigned char *As = Bs + g.b_chunk_bytes();                    // two X tiles [16][128][8] bf16
  unsigned char *Ys = As + 2 * X_BYTES;                          // Y tile, 3 splits of [NN/8][128][8] bf16
  uint4 *xtab = reinterpret_cast<uint4 *>(Ys + 3 * y_split_bytes);  // 256 x 16 B: bf16 images of 8 spins
  float *wg = reinterpret_cast<float *>(xtab + 256);               // NN: row idx of W (gate flip)
  float *corr = wg + NN;                                         // NN: correction of row idx
  cx<float> *wts = reinterpret_cast<cx<float> *>(corr + NN);     // [2 tiles][3][128]: w_sig, w_eta, conj(rho1)
  __shared__ __align__(8) uint64_t mbar[4];                      // GEMM 1 done, GEMM 2 done, X tile written, Y written
  __shared__ uint32_t tmem_base_sh;
  __shared__ __align__(16) uint4 xtail[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane;
  const int half = warp >> 2;  // column part of this warp (0 .. NT / 128 - 1)
  const uint32_t mbar1 = smem_u32(&mbar[0]), mbar2 = smem_u32(&mbar[1]);
  const uint32_t mbar_x = smem_u32(&mbar[2]), mbar_y = smem_u32(&mbar[3]);
  constexpr int NTB = NT + 32;  // threads of the CTA
  const bool mma_warp = warp == NT / 32;
  const int chunk = blockIdx.x;
  const int N = g.N, M = g.M, cols = 2 * M;
  const int W32 = (N + 31) >> 5;
  const bool gate = a.op2.kind == NKFB_OP_GATE;
  const int gidx = a.op2.idx;
  const float ssum = a.s0 + a.s1;
  const double F = a.weights ? 0.0 : a.sums[0] / a.sums[4];
  const float cvc = a.has_cv ? a.cv : 0.f;

  if (tid == 0) {
    mbar_init(mbar1, 1);
    mbar_init(mbar2, 1);
    mbar_init(mbar_x, NT / 32);
    mbar_init(mbar_y, NT / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_sh), 512);
  {  // stage the chunk of W~ (resident for the whole kernel), the gate row, zero the correction
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(a.prep) +
                                                       (size_t)chunk * g.b_chunk_bytes());
    uint4 *dst = reinterpret_cast<uint4 *>(Bs);
    const int n16 = (int)(g.b_chunk_bytes() / 16);
    for (int i = tid; i < n16; i += NTB) dst[i] = __ldg(src + i);
    for (int n = tid; n < NN; n += NTB) {
      const int col = chunk * NN + n;
      wg[n] = (gate && col < cols) ? psi.W[(int64_t)gidx * cols + col] : 0.f;
      corr[n] = 0.f;
    }
    // rows >= KC*8 of the X tiles (beyond the padded spin count) stay zero for the whole kernel
    uint4 *az = reinterpret_cast<uint4 *>(As);
    for (int i = tid; i < 2 * GX_CHUNKS * TC_ROWS; i += NTB) az[i] = make_uint4(0, 0, 0, 0);
    tc_fill_x_table(xtab, a.s0, a.s1);
    tc_fill_x_tail(xtail, N);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_theta = tmem_base_sh;        // two accumulators of NN columns: [0, NN) and [NN, 2 NN)
  const uint32_t tmem_g = tmem_base_sh + 2 * NN;   // NN columns: G of the current TC_G_FLUSH iterations (UMMA accumulator)
  const uint32_t tmem_gt = tmem_base_sh + 3 * NN;  // NN columns: G of all earlier iterations (added by the epilogue warps)
  const uint32_t idesc1 = umma_idesc_bf16_f32(TC_ROWS, NN, 0, 0);
  const uint32_t idesc2 = umma_idesc_bf16_f32(TC_ROWS, NN, 1, 1);
  uint32_t par1 = 0, par2 = 0;

  const int64_t n_tiles = (a.S + TC_ROWS - 1) / TC_ROWS;
  const int64_t t_begin = (int64_t)blockIdx.y * a.tiles_per_split;
  const int64_t t_end = min(n_tiles, t_begin + a.tiles_per_split);
  const int nh = NN / (NT / TC_ROWS);  // columns per part
  const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
  const int sshift = a.weights ? 0 : 1;  // sides per tile = 1 << sshift
  const int64_t K = t_end > t_begin ? (t_end - t_begin) << sshift : 0;
  const bool tail = (chunk + 1) * NN > cols;  // pseudo-unit / padding columns in this chunk
  float ccacc[NB8][8];
#pragma unroll
  for (int b = 0; b < NB8; ++b)
#pragma unroll
    for (int q = 0; q < 8; ++q) ccacc[b][q] = 0.f;

  // per-sample weights of a tile: w_eta = conj(A) + 2 c |A|^2, w_sig = 2 (L - F) - w_eta
  auto tile_weights = [&](int64_t tile) {
    if (tid < TC_ROWS) {
      const int64_t gs = tile * TC_ROWS + tid;
      cx<float> ws_ = mk<float>(0, 0), we_ = mk<float>(0, 0), rh_ = mk<float>(0, 0);
      if (gs < a.S && a.weights) {
        ws_ = a.weights[gs];
      } else if (gs < a.S) {
        const cx<float> A = cexp(a.log_val[gs]);
        const float A2 = abs2(A);
        const float Lv = A.re + cvc * (A2 - 1.f);
        we_ = mk<float>(A.re + 2.f * cvc * A2, -A.im);
        const float dl = (float)(2.0 * ((double)Lv - F));
        ws_ = mk<float>(dl - we_.re, -we_.im);
        rh_ = conj(a.rho1[gs]);
      }
      cx<float> *wt = wts + (tile & 1) * 3 * TC_ROWS;
      wt[tid] = ws_;
      wt[TC_ROWS + tid] = we_;
      wt[2 * TC_ROWS + tid] = rh_;
    }
  };
  auto issue_gemm1 = [&](int64_t k) {  // Theta[k & 1] = X[k & 1] W~   (one thread)
    const uint32_t a_base = smem_u32(As + (k & 1) * X_BYTES), b_base = smem_u32(Bs);
    const uint32_t td = tmem_theta + (uint32_t)((k & 1) * NN);
    for (int s = 0; s < 3; ++s)
      for (int ks = 0; ks < g.KS; ++ks) {
        const uint64_t ad = umma_desc(a_base + ks * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128);
        const uint64_t bd = umma_desc(b_base + s * (uint32_t)g.b_split_bytes() + ks * 2 * (NN * 16), NN * 16, 128);
        umma_bf16(td, ad, bd, idesc1, (s | ks) ? 1u : 0u);
      }
    umma_commit(mbar1);
  };

  auto issue_gemm2 = [&](int64_t k) {  // G += X[k & 1]^T Y   (one thread)
    const uint32_t a_base = smem_u32(As + (k & 1) * X_BYTES), y_base = smem_u32(Ys);
    for (int s = 0; s < 3; ++s)
      for (int ks = 0; ks < TC_ROWS / 16; ++ks) {
        // A: X tile read MN-major (rows = spins): 8-spin groups 2048 B apart (SBO), 8-sample groups 128 B apart (LBO)
        const uint64_t ad = umma_desc(a_base + ks * 256, 128, TC_ROWS * 16);
        // B: Y MN-major: 8-column groups 2048 B apart (SBO), 8-sample groups 128 B apart (LBO)
        const uint64_t bd = umma_desc(y_base + s * (uint32_t)y_split_bytes + ks * 256, 128, TC_ROWS * 16);
        umma_bf16(tmem_g, ad, bd, idesc2, ((k % TC_G_FLUSH) != 0 || s || ks) ? 1u : 0u);
      }
    umma_commit(mbar2);
  };
  // this warp's writes to shared memory (generic proxy) are done and visible to the tensor pipe: one arrive per warp
  auto warp_arrive = [&](uint32_t bar) {
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
  };

  if (mma_warp) {
    // ---------------------------------------------------------------- the issuing warp
    uint32_t px = 0, py = 0;
    for (int64_t k = 0; k < K; ++k) {
      if (k == 0) {
        mbar_wait_relaxed(mbar_x, px);  // X tile 0
        px ^= 1;
        tc_fence_after();
        if (lane == 0) issue_gemm1(0);
      }
      if (k + 1 < K) {
        mbar_wait_relaxed(mbar_x, px);  // X tile k + 1 (and every warp is past epilogue k - 1: Theta[(k + 1) & 1] is free)
        px ^= 1;
        tc_fence_after();
        if (lane == 0) issue_gemm1(k + 1);
      }
      mbar_wait_relaxed(mbar_y, py);  // Y(k)
      py ^= 1;
      tc_fence_after();
      if (lane == 0) issue_gemm2(k);
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- the epilogue warps
    if (K > 0) {  // weights of the first tile, X tile 0
      tile_weights(t_begin);
      tc_build_x_tile_fast<NT / TC_ROWS>(As, a.sigma, t_begin * TC_ROWS, a.S, g, xtab, xtail);
      warp_arrive(mbar_x);
      // the weights of the first tile are read by every warp: one barrier among the epilogue warps (named barrier 1)
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    }
    for (int64_t k = 0; k < K; ++k) {
      const int64_t tile = t_begin + (k >> sshift);
      const int side = (int)(k & ((1 << sshift) - 1));
      const int64_t s_base = tile * TC_ROWS;
      const uint32_t *packed = side == 0 ? a.sigma : a.eta;
      const bool flip = gate && side == 1;
      const cx<float> *wt = wts + (tile & 1) * 3 * TC_ROWS;
      const uint32_t th = tmem_theta + (uint32_t)((k & 1) * NN) + lane_off;

      mbar_wait(mbar1, par1);  // GEMM 1(k) done
      par1 ^= 1;
      tc_fence_after();
      const cx<float> w = side == 0 ? wt[row] : wt[TC_ROWS + row];
      const cx<float> r1 = flip ? wt[2 * TC_ROWS + row] : mk<float>(0, 0);  // conj(rho_1)
      const cx<float> r0 = mk<float>(1.f - r1.re, -r1.im);                    // conj(rho_0)
      float d = 0.f;
      if (flip) {
        const int64_t gs = s_base + row;
        if (gs < a.S) {
          const uint32_t wd = packed[gs * W32 + (gidx >> 5)];
          d = ssum - 2.f * (((wd >> (gidx & 31)) & 1u) ? a.s1 : a.s0);
        }
      }
      // Y[s, f] for this thread's row and block b8 of its part of the columns, as three bf16 splits
      auto compute = [&](int b8, GradEpi &o, auto tail_c) {
        constexpr bool TAIL = decltype(tail_c)::value;
        const int c0 = half * nh + 8 * b8;
        float v[8], y[8];
        tmem_ld8(th + (uint32_t)c0, v);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          cx<float> out = mk<float>(0, 0), c1 = mk<float>(0, 0);
          const int j = TAIL ? (chunk * NN + c0) / 2 + u : 0;  // hidden-unit index of this complex column
          if (!TAIL || j <= M) {
            cx<float> t0 = mk<float>(1, 0), t1 = mk<float>(1, 0);  // pseudo-unit j == M: O_{a_i} = x_i
            if (!TAIL || j < M) {
              t0 = conj(ctanh_fast(v[2 * u], v[2 * u + 1]));
              if (flip)
                t1 = conj(ctanh_fast(fmaf(d, wg[c0 + 2 * u], v[2 * u]), fmaf(d, wg[c0 + 2 * u + 1], v[2 * u + 1])));
            }
            if (flip) {
              c1 = w * (r1 * t1);
              out = w * (r0 * t0) + c1;
            } else {
              out = w * t0;
            }
          }
          y[2 * u] = out.re;
          y[2 * u + 1] = out.im;
          if (flip) {  // correction of row idx: sum over the samples of d_s w_eta conj(rho_1) conj(tanh theta')
            ccacc[b8][2 * u] = fmaf(d, c1.re, ccacc[b8][2 * u]);
            ccacc[b8][2 * u + 1] = fmaf(d, c1.im, ccacc[b8][2 * u + 1]);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) split3(y[2 * q], y[2 * q + 1], o.hi[q], o.mid[q], o.lo[q]);
      };
      // -> [c0/8][row][8] (MN-major B operand of GEMM 2)
      auto store = [&](int b8, const GradEpi &o) {
        const int c0 = half * nh + 8 * b8;
        unsigned char *dst = Ys + ((size_t)(c0 / 8) * TC_ROWS + row) * 16;
        *reinterpret_cast<uint4 *>(dst) = make_uint4(o.hi[0], o.hi[1], o.hi[2], o.hi[3]);
        *reinterpret_cast<uint4 *>(dst + y_split_bytes) = make_uint4(o.mid[0], o.mid[1], o.mid[2], o.mid[3]);
        *reinterpret_cast<uint4 *>(dst + 2 * y_split_bytes) = make_uint4(o.lo[0], o.lo[1], o.lo[2], o.lo[3]);
      };

      const bool chain_end = k > 0 && (k % TC_G_FLUSH) == 0;
      if (chain_end) {  // a chain of UMMAs has ended with GEMM 2(k - 1): G_total (+)= G_chain, this thread's row and columns
        mbar_wait(mbar2, par2);
        par2 ^= 1;
        tc_fence_after();
#pragma unroll
        for (int b8 = 0; b8 < NB8; ++b8) {
          const uint32_t c0 = (uint32_t)(half * nh + 8 * b8);
          float vs[8], vt[8];
          tmem_ld8(tmem_g + lane_off + c0, vs);
          if (k > TC_G_FLUSH) {
            tmem_ld8(tmem_gt + lane_off + c0, vt);
#pragma unroll
            for (int q = 0; q < 8; ++q) vs[q] += vt[q];
          }
          tmem_st8(tmem_gt + lane_off + c0, vs);
        }
      }
      GradEpi e0;
      if (tail)
        compute(0, e0, std::true_type{});
      else
        compute(0, e0, std::false_type{});
      if (k > 0 && !chain_end) {  // GEMM 2(k - 1) done: the Y buffer and X[(k + 1) & 1] are free
        mbar_wait(mbar2, par2);
        par2 ^= 1;
        tc_fence_after();
      }
      store(0, e0);
      if (k + 1 < K) {  // weights (first side of a new tile) and X tile of iteration k + 1
        const int64_t ntile = t_begin + ((k + 1) >> sshift);
        const int nside = (int)((k + 1) & ((1 << sshift) - 1));
        if (nside == 0) tile_weights(ntile);
        tc_build_x_tile_fast<NT / TC_ROWS>(As + ((k + 1) & 1) * X_BYTES, nside == 0 ? a.sigma : a.eta, ntile * TC_ROWS, a.S,
                                           g, xtab, xtail);
        warp_arrive(mbar_x);
      }
#pragma unroll
      for (int b8 = 1; b8 < NB8; ++b8) {
        GradEpi e;
        if (tail)
          compute(b8, e, std::true_type{});
        else
          compute(b8, e, std::false_type{});
        store(b8, e);
      }
      warp_arrive(mbar_y);
    }
    if (K > 0) {  // the last GEMM 2
      mbar_wait(mbar2, par2);
      tc_fence_after();
    }
  }

  // ---- correction of the gate row: reduce the per-thread sums over the 128 sample rows
  if (gate && !mma_warp) {
#pragma unroll
    for (int b8 = 0; b8 < NB8; ++b8)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float r = ccacc[b8][q];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
        if (lane == 0) atomicAdd(corr + half * nh + 8 * b8 + q, r);
      }
  }
  // ---- flush G (TMEM lane = gradient row: spins, then the bias row N) to the fp64 accumulators
  __syncthreads();
  if (K > 0 && !mma_warp) {
    const int i = row;
    const int64_t offK = 2 * (int64_t)M, offA = 2 * ((int64_t)M + (int64_t)N * M);
    for (int c0 = half * nh; c0 < (half + 1) * nh; c0 += 8) {
      float v[8];
      tmem_ld8(tmem_g + lane_off + (uint32_t)c0, v);  // the last (possibly short) chain
      if (K > TC_G_FLUSH) {
        float vt[8];
        tmem_ld8(tmem_gt + lane_off + (uint32_t)c0, vt);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += vt[q];
      }
      if (i <= N) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = chunk * NN + c0 + q;
          const int j = col >> 1, comp = col & 1;
          float val = v[q];
          if (gate && i == gidx) val += corr[c0 + q];
          if (j < M) {
            if (i < N)
              atomicAdd(a.grad_acc + offK + 2 * ((int64_t)i * M + j) + comp, (double)val);
            else if (a.has_bias)
              atomicAdd(a.grad_acc + 2 * j + comp, (double)val);
          } else if (j == M && i < N && a.has_vbias) {
            atomicAdd(a.grad_acc + offA + 2 * i + comp, (double)val);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_sh, 512);
}

// geometry of the tensor-core gradient for a network; false when the tensor-core path does not apply to it
static bool tc_grad_geometry(const nkfb_rbm_t *psi, TcGeom *gp, bool *wide_out, size_t *smem_out) {
  if (psi->dtype != NKFB_F32) return false;
  const int N = psi->n_visible, M = psi->n_hidden;
  if (N + 1 > TC_ROWS) return false;  // GEMM 2 holds all gradient rows (spins + bias) in one UMMA M = 128 tile
  TcGeom &g = *gp;
  // per column: three bf16 splits of a 128-sample Y tile (768 B) + gate row + correction (8 B); fixed: two X tiles, the
  // spin-group table, two tiles of per-sample weights.  Four accumulators of NN columns share the 512 TMEM columns.
  const size_t fixed = 2 * (size_t)GX_CHUNKS * TC_ROWS * 16 + 4096 + 2 * 3 * TC_ROWS * sizeof(cx<float>);
  auto smem_of = [&](const TcGeom &q) {
    return q.b_chunk_bytes() + fixed + 3 * (size_t)(q.NN / 8) * TC_ROWS * 16 + 2 * (size_t)q.NN * 4 + 128;
  };
  // sixteen warps (chunks of a multiple of 32 columns, up to 128) when that geometry fits; else eight (multiples of 16, up to 128: four accumulators of NN columns share the 512 TMEM columns)
  bool wide = !getenv("NKFB_GRAD_TC_NARROW") &&
              tc_geometry(N, M, &g, 2 * (M + 1), 3 * TC_ROWS * 2 + 8, fixed, 226 * 1024, 32, 128) && smem_of(g) <= (size_t)226 * 1024;
  if (!wide && !tc_geometry(N, M, &g, 2 * (M + 1), 3 * TC_ROWS * 2 + 8, fixed, 226 * 1024, 16, 128)) return false;
  const size_t smem = smem_of(g);
  if (smem > 227 * 1024) return false;
  *wide_out = wide;
  *smem_out = smem;
  return true;
}

// The bf16 splits of the network (the handle's third workspace); see tc_fwd_tables (forward_tc.cu) for `force` and
// NKFB_OPT_TC_TABLES_VALID.
static int tc_grad_tables(nkfb_handle_t h, const nkfb_rbm_t *psi, const TcGeom &g, bool force, cudaStream_t st) {
  const size_t need = g.prep_bytes_per_net();
  if (h->ws3_bytes < need) {
    if (h->ws3) cudaFree(h->ws3);
    h->ws3 = nullptr;
    h->ws3_bytes = 0;
    h->tc_grad_key.w[0] = nullptr;
    NKFB_CHECK_CUDA(h, cudaMalloc(&h->ws3, need));
    h->ws3_bytes = need;
  }
  nkfb_handle_s::TcTablesKey key;
  key.w[0] = key.w[1] = psi->kernel;
  key.b[0] = key.b[1] = psi->bias;
  key.n_visible = psi->n_visible;
  key.n_hidden = psi->n_hidden;
  key.NN = g.NN;
  key.NC = g.NC;
  key.KC = g.KC;
  if (!force && h->opt_tc_tables_valid && tc_key_equal(h->tc_grad_key, key)) return NKFB_OK;
  const int64_t tot = (int64_t)g.NC * g.KC * g.NN;
  const int grid = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 8);
  tc_prep_kernel<<<grid, 256, 0, st>>>(make_netdev<float>(psi), g, (__nv_bfloat16 *)h->ws3);
  NKFB_LAUNCHED(h);
  h->tc_grad_key = key;
  return NKFB_OK;
}

// nkfb_tc_prepare, gradient half: returns NKFB_OK (also when the tensor-core path does not apply: nothing to prepare)
int tc_grad_prepare(nkfb_handle_t h, const nkfb_rbm_t *psi, const double ls[2], cudaStream_t st) {
  TcGeom g;
  bool wide;
  size_t smem;
  if (!bf16_exact(ls[0]) || !bf16_exact(ls[1]) || !tc_grad_geometry(psi, &g, &wide, &smem)) return NKFB_OK;
  return tc_grad_tables(h, psi, g, true, st);
}

// returns NKFB_OK, an error, or 1 when the tensor-core path does not apply
int tc_infidelity_grad(nkfb_handle_t h, const nkfb_rbm_t *psi, const nkfb_op_t *op2, const uint32_t *sigma,
                       const uint32_t *eta, int64_t S, const double ls[2], double cv, const void *log_val,
                       const void *rho1, const double *sums, double *grad_acc, cudaStream_t st, const void *weights) {
  if (psi->dtype != NKFB_F32) return 1;
  if (op2 && op2->kind == NKFB_OP_ISING) return 1;
  if (!bf16_exact(ls[0]) || !bf16_exact(ls[1])) return 1;
  if (S < 2048) return 1;
  TcGeom g;
  bool wide;
  size_t smem;
  if (!tc_grad_geometry(psi, &g, &wide, &smem)) return 1;
  {
    const int rc = tc_grad_tables(h, psi, g, false, st);
    if (rc != NKFB_OK) return rc;
  }
  TcGradArgs a;
  a.op2 = make_opdev(op2);
  a.sigma = sigma;
  a.eta = eta;
  a.S = S;
  a.s0 = (float)ls[0];
  a.s1 = (float)ls[1];
  a.has_cv = !isnan(cv);
  a.cv = a.has_cv ? (float)cv : 0.f;
  a.log_val = (const cx<float> *)log_val;
  a.rho1 = (const cx<float> *)rho1;
  a.sums = sums;
  a.weights = (const cx<float> *)weights;
  a.grad_acc = grad_acc;
  a.prep = (const __nv_bfloat16 *)h->ws3;
  a.g = g;
  a.has_bias = psi->bias != nullptr;
  a.has_vbias = psi->visible_bias != nullptr;
  const int64_t tiles = (S + TC_ROWS - 1) / TC_ROWS;
  int64_t n_splits = std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)h->sm_count / g.NC));
  a.tiles_per_split = (tiles + n_splits - 1) / n_splits;
  n_splits = (tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  dim3 grid(g.NC, (unsigned)n_splits);
#define NKFB_GRAD_TC_CASE(NB, NT)                                                                                 \
  case NB:                                                                                                       \
    NKFB_CHECK_CUDA(h, cudaFuncSetAttribute(infid_grad_tc_kernel<NB, NT>,                                         \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    infid_grad_tc_kernel<NB, NT><<<grid, NT + 32, smem, st>>>(make_netdev<float>(psi), a);                           \
    break;
  if (wide) {
    switch (g.NN / 32) {
      NKFB_GRAD_TC_CASE(1, 512) NKFB_GRAD_TC_CASE(2, 512) NKFB_GRAD_TC_CASE(3, 512) NKFB_GRAD_TC_CASE(4, 512)
      default: return 1;
    }
  } else {
    switch (g.NN / 16) {
      NKFB_GRAD_TC_CASE(1, 256) NKFB_GRAD_TC_CASE(2, 256) NKFB_GRAD_TC_CASE(3, 256) NKFB_GRAD_TC_CASE(4, 256)
      NKFB_GRAD_TC_CASE(5, 256) NKFB_GRAD_TC_CASE(6, 256) NKFB_GRAD_TC_CASE(7, 256) NKFB_GRAD_TC_CASE(8, 256)
      default: return 1;  // wider chunks: CUDA-core kernel
    }
  }
#undef NKFB_GRAD_TC_CASE
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
