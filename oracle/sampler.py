"""Oracle: Metropolis-local samplers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

1. `metropolis_local_netket`: NetKet `nk.sampler.MetropolisLocal` semantics
   [NetKet 3.10; call sites infidelity/overlap/expect.py:26-27,
   README.md:55-58]: every chain draws its own site, the proposal flips it, the new
   log-probability comes from a FULL forward pass, accept iff u < exp(dlogp).  NumPy RNG.
   This is the algorithm the CPU baseline times.
2. `metropolis_local_b200`: line-by-line restatement (float64) of the device algorithm
   of netket_fidelity_b200/csrc/sampler.cu -- Philox4x32-10 counters, shared or
   per-chain site sequence, incremental theta update, log2-domain acceptance --
   used to check the CUDA kernels trajectory-by-trajectory.
"""
import numpy as np

from .philox import philox4x32_10
from .rbm import RBMParams, make_logfun

TAG_UNIFORM, TAG_SITE_CHAIN, TAG_SITE_SHARED, TAG_INIT = 0, 1, 2, 3
LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


def metropolis_local_netket(logfun, N, n_chains, chain_length, n_discard, sweep_size=None, rng=None,
                            init=None, local_states=(-1.0, 1.0), machine_pow=2):
    """Returns samples (n_chains, chain_length, N) and the final chain state."""
    rng = np.random.default_rng(0) if rng is None else rng
    sweep_size = N if sweep_size is None else sweep_size
    ls = np.asarray(local_states, dtype=np.float64)
    x = ls[rng.integers(0, 2, size=(n_chains, N))] if init is None else np.array(init, dtype=np.float64)
    logp = machine_pow * logfun(x).real
    out = np.empty((n_chains, chain_length, N))
    ar = np.arange(n_chains)
    for sweep in range(n_discard + chain_length):
        for _ in range(sweep_size):
            sites = rng.integers(0, N, size=n_chains)
            xp = x.copy()
            cur = xp[ar, sites]
            xp[ar, sites] = np.where(cur == ls[0], ls[1], ls[0])
            logp_new = machine_pow * logfun(xp).real          # full forward, as NetKet does
            acc = rng.random(n_chains) < np.exp(logp_new - logp)
            x[acc] = xp[acc]
            logp = np.where(acc, logp_new, logp)
        if sweep >= n_discard:
            out[:, sweep - n_discard] = x
    return out, x


def _bits_init(N, chain_id, key):
    W32 = (N + 31) // 32
    words = np.zeros(W32, dtype=np.uint32)
    for w in range(W32):
        r = philox4x32_10(np.array([w >> 2, 0, chain_id, TAG_INIT], dtype=np.uint32), key)
        v = int(r[w & 3])
        nb = N - 32 * w
        if nb < 32:
            v &= (1 << nb) - 1
        words[w] = v
    return words


def _get_bit(words, i):
    return (int(words[i >> 5]) >> (i & 31)) & 1


def _re_logcosh2_sum(th):
    """sum_j [|x| log2e + 1/2 log2(1 + t^2 + 2 t cos 2y)], t = exp(-2|x|)   (log2|cosh| + 1 per unit)."""
    ax = np.abs(th.real)
    t = np.exp(-2.0 * ax)
    return float(np.sum(ax * LOG2E + 0.5 * np.log2(1.0 + t * (2.0 * np.cos(2.0 * th.imag) + t))))


def metropolis_local_b200(p: RBMParams, n_chains, chain_length, n_discard, sweep_size, seed, chain_offset=0,
                          step_offset=0, site_mode=0, init_words=None, local_states=(-1.0, 1.0), U=None,
                          u_bits=32):
    """Device algorithm in float64.  Returns (samples words (n_chains, chain_length, W32) uint32,
    final words, n_accepted).  `U` = None for the plain target, or an oracle operator for the
    composite target |sum_c mel_c psi(x'_c)|^2 (accept test in natural-log units, as on the device)."""
    N, M = p.N, p.M
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    s0, s1 = local_states
    W32 = (N + 31) // 32
    Wk = p.kernel.astype(np.complex128)
    b = np.zeros(M, dtype=np.complex128) if p.bias is None else p.bias.astype(np.complex128)
    a = None if p.visible_bias is None else p.visible_bias.astype(np.complex128)
    logfun = make_logfun(p, U) if U is not None else None
    out = np.zeros((n_chains, chain_length, W32), dtype=np.uint32)
    fin = np.zeros((n_chains, W32), dtype=np.uint32)
    n_acc = 0
    for ch in range(n_chains):
        cid = chain_offset + ch
        words = _bits_init(N, cid, key) if init_words is None else np.array(init_words[ch], dtype=np.uint32)
        x = np.array([s1 if _get_bit(words, i) else s0 for i in range(N)], dtype=np.float64)
        th = b + x @ Wk
        if U is None:
            lc = _re_logcosh2_sum(th)
        else:
            logp = 2.0 * float(logfun(x[None, :])[0].real)
        step = step_offset
        for sweep in range(n_discard + chain_length):
            for _ in range(sweep_size):
                blk, comp = step >> 2, step & 3
                ru = philox4x32_10(np.array([blk & 0xFFFFFFFF, blk >> 32, cid, TAG_UNIFORM], dtype=np.uint32), key)
                if site_mode == 0:
                    rs = philox4x32_10(np.array([blk & 0xFFFFFFFF, blk >> 32, 0, TAG_SITE_SHARED], dtype=np.uint32), key)
                else:
                    rs = philox4x32_10(np.array([blk & 0xFFFFFFFF, blk >> 32, cid, TAG_SITE_CHAIN], dtype=np.uint32), key)
                site = (int(rs[comp]) * N) >> 32
                r = int(ru[comp])
                if u_bits == 32:
                    log2u = np.log2((r + 0.5) / 4294967296.0)
                else:
                    log2u = np.log2(((r >> 8) + 0.5) / 16777216.0)
                xold = x[site]
                d = (s0 + s1) - 2.0 * xold
                if U is None:
                    thn = th + d * Wk[site]
                    lcn = _re_logcosh2_sum(thn)
                    dl = 2.0 * (lcn - lc) + 2.0 * LOG2E * d * (a[site].real if a is not None else 0.0)
                    acc = log2u < dl
                    if acc:
                        th, lc = thn, lcn
                else:
                    xn = x.copy()
                    xn[site] = xold + d
                    logpn = 2.0 * float(logfun(xn[None, :])[0].real)
                    acc = log2u * LN2 < logpn - logp
                    if acc:
                        logp = logpn
                if acc:
                    x[site] = xold + d
                    words[site >> 5] ^= np.uint32(1 << (site & 31))
                    n_acc += 1
                step += 1
            if sweep >= n_discard:
                out[ch, sweep - n_discard] = words
        fin[ch] = words
    return out, fin, n_acc
