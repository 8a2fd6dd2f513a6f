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
struct LcAcc {
  float sx, sy, lr, li;  // sum |x|, sum of the folded y, accumulated log of the flushed products
  __device__ __forceinline__ void reset() { sx = sy = lr = li = 0.f; }
  // log cosh z = |x| + log((1 + t e^{-2iy}) / 2) (folded): multiply the running product by (1 + t e^{-2iy}) / 2,
  // which is ~1 for small theta, so that the log taken every 8 units is small and keeps its absolute accuracy
  __device__ __forceinline__ void add(float x, float y, float &pr, float &pi) {
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
  __device__ __forceinline__ void flush(float &pr, float &pi) {
    lr += 0.5f * logf(fmaf(pr, pr, pi * pi));
    li += atan2f(pi, pr);
    pr = 1.f;
    pi = 0.f;
  }
  __device__ __forceinline__ cx<float> value() const { return mk<float>(sx + lr, sy + li); }
};

// Epilogue of one (chunk, tile, side): this thread's row, columns [c_begin, c_end) of the chunk (multiples of 8).
// FLIP: also accumulate the configuration with the gate site flipped (theta + d W[idx, :]);  FULL: every column of
// the range is a real hidden-unit column (no per-unit bound test; false only for the last chunk).
// Blocks of 8 units (16 columns) with one flush of the running products each, as before (same numerics).
template <bool FLIP, bool FULL>
__device__ __forceinline__ void fwd_epilogue(LcAcc &A0, LcAcc &A1, uint32_t trow, int c_begin, int c_end, int col0,
                                             int cols, const float *wg, float d) {
  float p0r = 1.f, p0i = 0.f, p1r = 1.f, p1i = 0.f;
  int c0 = c_begin;
#pragma unroll 1
  for (; c0 + 16 <= c_end; c0 += 16) {
    float v[16];
    tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (FULL || col0 + c0 + 2 * u < cols) {
        A0.add(v[2 * u], v[2 * u + 1], p0r, p0i);
        if (FLIP) A1.add(fmaf(d, wg[c0 + 2 * u], v[2 * u]), fmaf(d, wg[c0 + 2 * u + 1], v[2 * u + 1]), p1r, p1i);
      }
    }
    A0.flush(p0r, p0i);
    if (FLIP) A1.flush(p1r, p1i);
  }
  if (c0 < c_end) {  // a last block of 8 columns
    float v[8];
    tmem_ld8(trow + (uint32_t)c0, v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (FULL || col0 + c0 + 2 * u < cols) {
        A0.add(v[2 * u], v[2 * u + 1], p0r, p0i);
        if (FLIP) A1.add(fmaf(d, wg[c0 + 2 * u], v[2 * u]), fmaf(d, wg[c0 + 2 * u + 1], v[2 * u + 1]), p1r, p1i);
      }
    }
    A0.flush(p0r, p0i);
    if (FLIP) A1.flush(p1r, p1i);
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
  int single_load;  // X tile: one packed-row load per thread (tc_build_x_tile)
  int tmem_cols;    // TMEM columns this CTA allocates
};

__global__ void __launch_bounds__(TC_THREADS, 2) infid_fwd_tc_kernel(NetDev<float> psi, NetDev<float> phi, TcFwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcGeom g = a.g;
  unsigned char *Bs = smem;                             // 3 splits of the current chunk
  unsigned char *As = Bs + g.b_chunk_bytes();           // X tile
  uint4 *xtab = reinterpret_cast<uint4 *>(As + g.a_bytes());  // 256 x 16 B: bf16 images of 8 spins (tc_build_x_tile)
  float *wg = reinterpret_cast<float *>(xtab + 256);        // [2][NN] floats: row `idx` of W (this chunk) for the gate flip
  cx<float> *xch = reinterpret_cast<cx<float> *>(wg + 2 * g.NN);  // exchange buffer [7][TC_ROWS]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_sh;
  __shared__ double red[8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane;  // TMEM lane == sample row of the tile
  const int half = warp >> 2;              // which half of the chunk's columns this thread reduces
  const uint32_t mbar_a = smem_u32(&mbar);

  if (tid == 0) {
    mbar_init(mbar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // 256 columns per CTA (two CTAs share the SM's 512); all 512 when the CTA is alone on its SM (a.tmem_cols)
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_sh), a.tmem_cols);
  tc_fill_x_table(xtab, a.s0, a.s1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;
  const uint32_t idesc = umma_idesc_bf16_f32(TC_ROWS, g.NN, 0, 0);
  uint32_t parity = 0;

  const float ssum = a.s0 + a.s1;
  const int W32 = (g.N + 31) >> 5;
  const int cols = 2 * g.M;
  double sL = 0, sL2 = 0, sA2 = 0, sRe = 0, sn = 0, sbad = 0;  // sbad: local estimators that are NaN or infinite

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
      const NetDev<float> &net = netid == 0 ? psi : phi;
      const __nv_bfloat16 *prep = netid == 0 ? a.prep_psi : a.prep_phi;
#pragma unroll 1
      for (int chunk = 0; chunk < g.NC; ++chunk) {
        __syncthreads();  // everyone is done with the previous chunk's Bs / wg
        {                 // stage the three bf16 splits of this chunk
          const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(prep) +
                                                             (size_t)chunk * g.b_chunk_bytes());
          uint4 *dst = reinterpret_cast<uint4 *>(Bs);
          const int n16 = (int)(g.b_chunk_bytes() / 16);
          for (int i = tid; i < n16; i += TC_THREADS) dst[i] = __ldg(src + i);
        }
        // The 2 TC_T phases (tile, side) of this chunk, software-pipelined when two accumulators fit the CTA's 256
        // TMEM columns: the X tile and the UMMAs of phase p + 1 are issued BEFORE the epilogue of phase p (the X
        // buffer is free as soon as the UMMAs of phase p have completed), so the tensor pipe works under the
        // epilogue instead of in front of it.
        const bool dbuf = 2 * g.NN <= a.tmem_cols;
        auto issue = [&](int p) {
          const int t = p >> 1, side = p & 1;
          const int ev = netid * 2 + side;
          const OpDev &op = (ev == 1) ? a.op2 : (ev == 2 ? a.op1 : a.op4);
          const bool flip = (ev != 0) && (op.kind == NKFB_OP_GATE);
          const uint32_t *packed = side == 0 ? a.sigma : a.eta;
          const int64_t s_base = (grp * TC_T + t) * TC_ROWS;
          if (flip) {
            float *wgp = wg + (p & 1) * g.NN;
            for (int n = tid; n < g.NN; n += TC_THREADS) {
              const int col = chunk * g.NN + n;
              wgp[n] = col < cols ? net.W[(int64_t)op.idx * cols + col] : 0.f;
            }
          }
          tc_build_x_tile(As, packed, s_base, a.S, g, a.s0, a.s1, a.single_load != 0, xtab);
          fence_async_smem();
          tc_fence_before();
          __syncthreads();
          if (tid == 0) {
            tc_fence_after();
            const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs);
            const uint32_t td = tmem_d + (dbuf ? (uint32_t)((p & 1) * g.NN) : 0u);
            for (int s = 0; s < 3; ++s)
              for (int ks = 0; ks < g.KS; ++ks) {
                const uint64_t ad = umma_desc(a_base + ks * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128);
                const uint64_t bd = umma_desc(b_base + s * (uint32_t)g.b_split_bytes() + ks * 2 * (g.NN * 16),
                                              g.NN * 16, 128);
                umma_bf16(td, ad, bd, idesc, (s | ks) ? 1u : 0u);
              }
            umma_commit(mbar_a);
          }
        };
        if (dbuf) issue(0);
#pragma unroll
        for (int p = 0; p < 2 * TC_T; ++p) {
          const int t = p >> 1, side = p & 1;
          const int ev = netid * 2 + side;
          const OpDev &op = (ev == 1) ? a.op2 : (ev == 2 ? a.op1 : a.op4);
          const bool flip = (ev != 0) && (op.kind == NKFB_OP_GATE);
          const uint32_t *packed = side == 0 ? a.sigma : a.eta;
          const int64_t s_base = (grp * TC_T + t) * TC_ROWS;
          if (!dbuf) issue(p);
          mbar_wait(mbar_a, parity);
          parity ^= 1;
          tc_fence_after();
          if (dbuf && p + 1 < 2 * TC_T) issue(p + 1);
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
          const int nh = g.NN / 2;  // columns per half (multiple of 8)
          const uint32_t trow = tmem_d + (dbuf ? (uint32_t)((p & 1) * g.NN) : 0u) + ((uint32_t)(32 * (warp & 3)) << 16);
          const float *wgp = wg + (p & 1) * g.NN;
          const int cb = half * nh, ce = (half + 1) * nh, col0 = chunk * g.NN;
          const bool full = col0 + ce <= cols;
          if (flip) {
            if (full)
              fwd_epilogue<true, true>(A0, A1, trow, cb, ce, col0, cols, wgp, d);
            else
              fwd_epilogue<true, false>(A0, A1, trow, cb, ce, col0, cols, wgp, d);
          } else {
            if (full)
              fwd_epilogue<false, true>(A0, A1, trow, cb, ce, col0, cols, wgp, d);
            else
              fwd_epilogue<false, false>(A0, A1, trow, cb, ce, col0, cols, wgp, d);
          }
          tc_fence_before();
          __syncthreads();  // this phase's TMEM buffer and gate row may be overwritten now
        }
      }
    }

    // ---- combine the two column halves, add the visible-bias terms, finish the estimator
#pragma unroll
    for (int t = 0; t < TC_T; ++t) {
      const int64_t s_base = (grp * TC_T + t) * TC_ROWS;
      __syncthreads();
      if (half == 1) {
        xch[0 * TC_ROWS + row] = acc0[t][0].value();
        xch[1 * TC_ROWS + row] = acc0[t][1].value();
        xch[2 * TC_ROWS + row] = acc0[t][2].value();
        xch[3 * TC_ROWS + row] = acc0[t][3].value();
        xch[4 * TC_ROWS + row] = acc1[t][1].value();
        xch[5 * TC_ROWS + row] = acc1[t][2].value();
        xch[6 * TC_ROWS + row] = acc1[t][3].value();
      }
      __syncthreads();
      const int64_t gs = s_base + row;
      if (half == 0 && gs < a.S) {
        cx<float> b0[4], b1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) b0[e] = acc0[t][e].value() + xch[e * TC_ROWS + row];
        b1[0] = mk<float>(0, 0);
#pragma unroll
        for (int e = 1; e < 4; ++e) b1[e] = acc1[t][e].value() + xch[(3 + e) * TC_ROWS + row];
        // visible-bias terms a.x for (psi, sigma) (psi, eta) (phi, sigma) (phi, eta)
        cx<float> ax[4] = {mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0), mk<float>(0, 0)};
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
        cx<float> T[4], rho = mk<float>(0, 0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const OpDev &op = (e == 1) ? a.op2 : (e == 2 ? a.op1 : a.op4);
          const NetDev<float> &net = e < 2 ? psi : phi;
          const uint32_t *packed = (e & 1) ? a.eta : a.sigma;
          const cx<float> l0 = b0[e] + ax[e];
          if (e == 0 || op.kind != NKFB_OP_GATE) {
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

  double v[6] = {sL, sL2, sA2, sRe, sn, sbad};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double r = block_sum_256(v[q], red);
    if (threadIdx.x == 0) atomicAdd(a.sums + q, r);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, a.tmem_cols);
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
  // two CTAs per SM (each its own TMEM accumulator): one CTA's UMMAs and X-tile builds overlap the other's
  // log-cosh epilogue; the W~ chunk of a CTA is sized for half of the shared memory
  size_t budget = 104 * 1024;
  if (const char *e = getenv("NKFB_TC_FWD_BUDGET_KB")) budget = (size_t)atoi(e) * 1024;
  const size_t xb = (size_t)(((psi->n_visible + 1 + 15) / 16) * 2) * TC_ROWS * 16 + 4096;  // X tile + spin-group table
  if (!tc_geometry(psi->n_visible, psi->n_hidden, &g, 2 * psi->n_hidden, 0, xb, budget) &&
      !tc_geometry(psi->n_visible, psi->n_hidden, &g, 2 * psi->n_hidden, 0, xb))
    return 1;
  const size_t smem = g.b_chunk_bytes() + g.a_bytes() + 4096 + (size_t)g.NN * 8 + 7 * TC_ROWS * sizeof(cx<float>) + 128;
  if (smem > 227 * 1024) return 1;

  // prepared bf16 splits of both networks live in the handle's second workspace
  const size_t need = 2 * g.prep_bytes_per_net();
  if (h->ws2_bytes < need) {
    if (h->ws2) cudaFree(h->ws2);
    h->ws2 = nullptr;
    h->ws2_bytes = 0;
    NKFB_CHECK_CUDA(h, cudaMalloc(&h->ws2, need));
    h->ws2_bytes = need;
  }
  __nv_bfloat16 *prep_psi = (__nv_bfloat16 *)h->ws2;
  __nv_bfloat16 *prep_phi = (__nv_bfloat16 *)((char *)h->ws2 + g.prep_bytes_per_net());
  {
    const int64_t tot = (int64_t)g.NC * g.KC * g.NN;
    const int grid = (int)std::min<int64_t>((tot + 255) / 256, (int64_t)h->sm_count * 8);
    tc_prep_kernel<<<grid, 256, 0, st>>>(make_netdev<float>(psi), g, prep_psi);
    NKFB_LAUNCHED(h);
    tc_prep_kernel<<<grid, 256, 0, st>>>(make_netdev<float>(phi), g, prep_phi);
    NKFB_LAUNCHED(h);
  }
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
  a.single_load = getenv("NKFB_TC_FWD_SINGLE") ? 1 : 0;
  a.tmem_cols = smem > (size_t)115 * 1024 ? 512 : 256;  // more than half of the SM's shared memory: one CTA per SM
  NKFB_CHECK_CUDA(h, cudaFuncSetAttribute(infid_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = smem <= 112 * 1024 ? 2 : 1;
  const int grid = (int)std::min<int64_t>(a.n_groups, (int64_t)h->sm_count * per_sm);
  infid_fwd_tc_kernel<<<grid, TC_THREADS, smem, st>>>(make_netdev<float>(psi), make_netdev<float>(phi), a);
  NKFB_LAUNCHED(h);
  return NKFB_OK;
}

}  // namespace nkfb
